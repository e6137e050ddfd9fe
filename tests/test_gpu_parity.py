"""Parity of the CUDA path (through the C ABI) with the reference: golden fixtures produced by
the reference itself, the oracle on seeded inputs, and the reference's own golden file."""
import os

import numpy as np
import pytest

from tests.golden_util import C1, need_c1, G1, G2, kat

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g1():
    from metacache_b200.database import Database
    g = G1()
    g.db = Database(0, 1)
    g.db.load_part_arrays(0, g.keys, g.sizes, g.values, batch=50000)   # several append batches
    yield g
    g.db.close()


def _sk(g):
    from metacache_b200.database import SketchingOpt
    return SketchingOpt(g.k, g.s, g.w, g.stride)


def test_kat_sketches_on_gpu(g1):
    from metacache_b200.database import query_reads
    z = kat()
    res = query_reads(g1.db, [z["S"].tobytes(), z["amb_seq"].tobytes(), z["low_seq"].tobytes()],
                      _sk(g1), with_sketches=True)
    assert np.array_equal(res[0][2][0], z["S_win0"]) and np.array_equal(res[0][2][1], z["S_win1"])
    assert np.array_equal(res[1][2][0], z["amb_feats"])
    assert np.array_equal(res[2][2][0], z["low_feats"])


def _check(res, exp, reads, k):
    for i, (allh, top, sks) in enumerate(res):
        # the GPU enumerates every window; the reference only sketches those with >= k characters
        a, b = reads[i]
        kept = [s for s, keep in zip(sks, _window_keeps(a, b, k)) if keep]
        assert len(kept) == len(exp.sketches[i]), i
        for x, y in zip(kept, exp.sketches[i]):
            assert np.array_equal(x, y), i
        assert np.array_equal(allh, exp.allhits[i]), i
        assert top == exp.top[i], i


def _window_keeps(a, b, k, w=127, stride=112):
    """for every window the GPU enumerates (mate 1 then mate 2, empty mate 2 not enumerated;
    an empty mate 1 with a non-empty mate 2 is dropped): True if it has >= k characters"""
    keeps = []
    seqs = [a] if len(b) == 0 else ([b] if len(a) == 0 else [a, b])
    for s in seqs:
        n = len(s)
        if n <= w:
            keeps.append(n >= k)
            continue
        full = (n - w) // stride + 1
        keeps += [True] * full
        if full * stride < n:
            keeps.append(n - full * stride >= k)
    return keeps


@pytest.mark.parametrize("tag,maxc,insert", [("c2_", 2, 0), ("c5_", 5, 0), ("c2i_", 2, 1000)])
def test_g1_matches_reference(g1, tag, maxc, insert):
    from metacache_b200.database import query_reads
    res = query_reads(g1.db, g1.reads, _sk(g1), max_candidates=maxc, insert_size_max=insert,
                      copy_all_hits=True, with_sketches=True, batch_queries=300)
    assert len(res) == len(g1.reads)
    _check(res, g1.expected(tag), g1.reads, g1.k)


def test_g1_tophits_without_allhits_and_tiny_batches(g1):
    from metacache_b200.database import query_reads
    exp = g1.expected("c2_")
    res = query_reads(g1.db, g1.reads, _sk(g1), copy_all_hits=False, batch_queries=7)
    assert [r[1] for r in res] == exp.top


def test_two_parts_merge_matches_reference_per_part():
    from metacache_b200.database import Database, query_reads
    from oracle import mc_oracle as O
    g1, g2 = G1(), G2()
    db = Database(0, 2)
    for p in (0, 1):
        db.load_part_arrays(p, *g2.parts[p])
    res = query_reads(db, g1.reads, copy_all_hits=True)
    e0, e1 = g2.expected(0), g2.expected(1)
    for i, (allh, top) in enumerate(res):
        assert np.array_equal(allh, np.concatenate([e0.allhits[i], e1.allhits[i]])), i
        assert top == O.merge_tops([e0.top[i], e1.top[i]], 2), i
    db.close()


def test_lowest_rank_merge_matches_oracle(g1):
    """`-lowest` above sequence: candidates of targets sharing a taxon are merged
    (candidate_generation.hpp:203-228)."""
    from metacache_b200.database import query_reads
    from oracle import mc_oracle as O
    ntgt = len(g1.target_windows)
    # G0,G0m1,G0m2 share taxon 100, G3* 101, G7* 102, target 5 has no ancestor (dropped)
    tax = np.arange(1, ntgt + 1, dtype=np.uint64)
    for base, key in ((0, 100), (3, 101), (7, 102)):
        tax[base] = key
    tax[10:12], tax[12:14], tax[14:16], tax[5] = 100, 101, 102, 0
    g1.db.set_target_taxa(tax)
    tab = O.Table(g1.keys, g1.sizes, g1.values)
    try:
        for maxc in (2, 4):
            res = query_reads(g1.db, g1.reads, _sk(g1), max_candidates=maxc, copy_all_hits=False)
            for i, (a, b) in enumerate(g1.reads):
                _, top = O.query(tab, a, b, maxc=maxc, tax_of_tgt=tax)
                assert res[i][1] == top, (i, maxc)
    finally:
        g1.db.set_target_taxa(None)


def test_heavy_paths_give_identical_results(g1):
    """force reads through the CTA kernel (shared and global-memory variants)"""
    import ctypes as C
    from metacache_b200 import _lib
    from metacache_b200.database import QueryBatch, make_candidate_generation_rules
    exp = g1.expected("c2_")
    heavy = [i for i, a in enumerate(exp.allhits) if len(a) > 512]
    assert len(heavy) >= 5            # the repeat target produces thousands of locations
    huge = [i for i, a in enumerate(exp.allhits) if len(a) > 16384]
    assert len(huge) >= 1             # exercises the global-scratch variant
    qb = QueryBatch(g1.db, 64, 1 << 20, 2, True, 1)
    hd = qb.host_data(0)
    for i in heavy[:60]:
        a, b = g1.reads[i]
        assert qb.add_paired_read(0, a, b, _sk(g1), make_candidate_generation_rules(len(a), len(b)))
    g1.db.query_gpu_async(qb, 0, _sk(g1))
    hd.wait_for_results()
    for s, i in enumerate(heavy[:60]):
        assert np.array_equal(hd.allhits(s), exp.allhits[i]), i
        assert [c.as_tuple() for c in hd.top_candidates(s) if c.hits] == exp.top[i], i
    qb.close()


def test_device_builder_matches_reference_build(g1):
    """mcb200_db_build_part_from_targets vs the reference `metacache build` of the same FASTA"""
    import ctypes as C
    import torch
    from metacache_b200 import _lib
    from metacache_b200.database import Database
    flat = np.concatenate([np.frombuffer(t, np.uint8) for t in g1.targets])
    off = np.zeros(len(g1.targets) + 1, np.uint64)
    off[1:] = np.cumsum([len(t) for t in g1.targets])
    d_bases = torch.from_numpy(flat).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    db = Database(0, 1)
    sk = _sk(g1).c()
    wins = np.zeros(len(g1.targets), np.uint32)
    _lib.check(_lib.lib().mcb200_db_build_part_from_targets(
        db._h, 0, d_bases.data_ptr(), d_off.data_ptr(), len(g1.targets), 0, C.byref(sk), 254, 0.0,
        wins.ctypes.data))
    assert np.array_equal(wins, g1.target_windows)
    keys, sizes, values = db.export_part(0)
    assert len(keys) == len(g1.keys) and len(values) == len(g1.values)
    o = np.argsort(keys)
    ro = np.argsort(g1.keys)
    assert np.array_equal(keys[o], g1.keys[ro]) and np.array_equal(sizes[o], g1.sizes[ro])
    offs = np.zeros(len(sizes) + 1, np.int64); np.cumsum(sizes, out=offs[1:])
    roffs = np.zeros(len(g1.sizes) + 1, np.int64); np.cumsum(g1.sizes, out=roffs[1:])
    for a, b in zip(o, ro):
        assert np.array_equal(values[offs[a]:offs[a + 1]], g1.values[roffs[b]:roffs[b + 1]])
    db.close()


def test_random_reads_against_oracle(g1):
    """seeded synthetic reads (C2-like recipe at small scale) vs the oracle"""
    from metacache_b200.database import query_reads
    from oracle import mc_oracle as O
    rng = np.random.default_rng(7)
    tab = O.Table(g1.keys, g1.sizes, g1.values)
    reads = []
    acgt = np.frombuffer(b"ACGT", np.uint8)
    for _ in range(3000):
        t = g1.targets[rng.integers(len(g1.targets))]
        o = int(rng.integers(0, len(t) - 150))
        a = np.frombuffer(t[o:o + 150], np.uint8).copy()
        m = rng.random(150) < 0.01
        a[m] = rng.choice(acgt, int(m.sum()))
        a[rng.random(150) < 0.001] = ord("N")
        reads.append(a.tobytes())
    res = query_reads(g1.db, reads, _sk(g1), copy_all_hits=True, batch_queries=1024)
    for i, r in enumerate(reads):
        allh, top = O.query(tab, r, b"")
        assert np.array_equal(res[i][0], allh), i
        assert res[i][1] == top, i


def test_c1_bundled_database_matches_reference_golden_file():
    """BASELINE config C1: reference-built `bacteria1` + bundled reads vs classified.expected"""
    need_c1(gpu_test=True)
    from metacache_b200 import formatting
    from metacache_b200.database import Database, query_reads
    from oracle import refio
    db = Database.read(os.path.join(C1, "bacteria1"))
    names = db.meta.target_names()
    expected, section = {}, None
    for line in open(os.path.join(C1, "classified.expected")):
        if line.startswith("# data/"):
            section = line.strip()[7:]
        if line.startswith("#"):
            continue
        cols = line.rstrip("\n").split("\t|\t")
        if len(cols) == 6 and cols[0].isdigit():
            expected.setdefault(section, {})[int(cols[0])] = (cols[3], cols[4], cols[5])   # by query id
    db.copy_target_lineages_to_gpus()
    single = refio.read_fasta(os.path.join(C1, "single.fa"))
    pf = refio.read_fasta(os.path.join(C1, "pairs.fa"))
    p1 = refio.read_fasta(os.path.join(C1, "pair.1.fa"))
    p2 = refio.read_fasta(os.path.join(C1, "pair.2.fa"))
    runs = {
        "single": [(s, b"") for _, s in single],
        "pairs": [(pf[i][1], pf[i + 1][1]) for i in range(0, len(pf), 2)],
        "pair.1 + data/pair.2": [(a[1], b[1]) for a, b in zip(p1, p2)],
    }
    from metacache_b200.statistics import ClassificationStatistics, TaxonCounts
    from tests.golden_util import reference_abundance_blocks
    blocks = reference_abundance_blocks(os.path.join(C1, "cli_cpu_reference.out"))
    assert len(blocks) == len(runs)
    total = classified = 0
    for si, (sec, items) in enumerate(runs.items()):
        exp = expected[sec]
        res = query_reads(db, items, copy_all_hits=True, classify=True)
        # -abundances -abundance-per species from the DEVICE's classifications (test/run_tests:153), against
        # what the unmodified CPU reference prints (estimate_abundance, classification.cpp:304-377)
        pairs = np.asarray([r[2] for r in res], np.uint32).reshape(-1, 2)
        st, tc = ClassificationStatistics(), TaxonCounts(db.meta.taxa)
        st.assign_batch(pairs)
        tc.count_batch(pairs)
        lines = tc.abundance_lines(st)
        tc.estimate_abundance(4)
        assert lines + tc.estimate_lines(st, 4) == blocks[si], sec
        for qid, (allh, top, cls) in enumerate(res, start=1):
            if qid not in exp:
                continue
            assert formatting.format_all_hits(allh, names) == exp[qid][0], (sec, qid)
            want = exp[qid][1] if exp[qid][1] != "--" else ""
            assert formatting.format_top_hits(top, names) == want, (sec, qid)
            # classification column (classify(): LCA of the candidates above the hit threshold)
            assert formatting.format_classification(cls[0], cls[1], db.meta.taxa) == exp[qid][2], (sec, qid)
            total += 1
            classified += cls[0] != 0
    assert total > 30000 and classified > 150
    db.close()


@pytest.mark.parametrize("maxc", [1, 2, 5])
def test_fast_kernel_tophits_only_matches_reference(g1, maxc):
    """top hits without all-hits run through the sort-free kernel (query_fast_kernel)"""
    from metacache_b200.database import query_reads
    from oracle import mc_oracle as O
    res = query_reads(g1.db, g1.reads, _sk(g1), max_candidates=maxc, copy_all_hits=False, batch_queries=500)
    if maxc in (2, 5):
        exp = g1.expected("c2_" if maxc == 2 else "c5_")
        assert [r[1] for r in res] == exp.top
    else:
        tab = O.Table(g1.keys, g1.sizes, g1.values)
        for i, (a, b) in enumerate(g1.reads):
            assert res[i][1] == O.query(tab, a, b, maxc=1)[1], i


def test_fast_kernel_small_table_overflows_to_cta_kernel(g1):
    """a 128-slot aggregation table sends more reads through the CTA kernel; results identical"""
    import ctypes as C
    from metacache_b200 import _lib
    from metacache_b200.database import QueryBatch, make_candidate_generation_rules
    exp = g1.expected("c2_")
    L = _lib.lib()
    import torch
    reads = g1.reads
    flat = np.concatenate([np.frombuffer(a + b, np.uint8) for a, b in reads] + [np.zeros(64, np.uint8)])
    seq_off, seq_qry, max_win = [0], [], []
    for i, (a, b) in enumerate(reads):
        for s in ((a, b) if len(b) and len(a) else ((a,) if len(a) or not len(b) else (b,))):
            seq_off.append(seq_off[-1] + len(s)); seq_qry.append(i)
        max_win.append(2 + (len(a) + len(b)) // 112)
    # note: concatenation order must match: rebuild flat in the same order
    chunks = []
    for a, b in reads:
        for s in ((a, b) if len(b) and len(a) else ((a,) if len(a) or not len(b) else (b,))):
            chunks.append(np.frombuffer(s, np.uint8))
    flat = np.concatenate(chunks + [np.zeros(64, np.uint8)])
    dev = torch.device("cuda", 0)
    t_b = torch.from_numpy(flat).to(dev)
    t_o = torch.tensor(seq_off, dtype=torch.int32, device=dev)
    t_q = torch.tensor(seq_qry, dtype=torch.int32, device=dev)
    t_w = torch.tensor(max_win, dtype=torch.int32, device=dev)
    nq, ns, nb = len(reads), len(seq_qry), seq_off[-1]
    ws = _lib.check_ptr(L.mcb200_workspace_create(g1.db._h, nq, ns, nb + 64, 2, 0))
    _lib.check(L.mcb200_workspace_set_warp_capacity(ws, 128))
    _lib.check(L.mcb200_workspace_set_profiling(ws, 1))
    q = _lib.DevQueries(t_b.data_ptr(), t_o.data_ptr(), t_q.data_ptr(), t_w.data_ptr(), ns, nq, nb)
    sk = _sk(g1).c()
    top = torch.empty((nq, 2, 4), dtype=torch.int32, device=dev)
    # top hits from the table: the 128-slot tables overflow for many reads (second pass, CTA tiers); the 19 kbp
    # tandem-repeat read of g1 (660 k locations, 254 distinct) is settled by the distinct-location CTA tier
    assert _lib.query_device_checked(ws, q, sk, top.data_ptr()) == 1
    cnt = (C.c_uint64 * 8)()
    _lib.check(L.mcb200_workspace_counters(ws, cnt))
    assert cnt[0] + cnt[1] + cnt[2] == nq and cnt[1] > 0

    def check_tops():
        got = top.cpu().numpy().astype(np.uint32)
        for i in range(nq):
            tl = [tuple(int(x) for x in row) for row in got[i] if row[1] > 0]
            assert tl == exp.top[i], i
    check_tops()
    # The order-dependent taxon merge (here with one taxon per target, i.e. the same result) sorts RAW location
    # lists: that read then needs more than a CTA's region of the default scratch pool.  The asynchronous
    # device API must not lose it silently: the check after the call reports
    # MCB200_EAGAIN (pool grown), the re-issued call is complete.
    g1.db.set_target_taxa(np.arange(1, len(g1.targets) + 1, dtype=np.uint64))
    try:
        top.zero_()
        _lib.check(L.mcb200_query_device(ws, C.byref(q), C.byref(sk), top.data_ptr(), None))
        assert L.mcb200_workspace_check(ws) == _lib.EAGAIN
        assert b"scratch pool grown" in L.mcb200_last_error()
        _lib.check(L.mcb200_workspace_counters(ws, cnt))             # resets the counters of the incomplete attempt
        assert _lib.query_device_checked(ws, q, sk, top.data_ptr()) == 1
        _lib.check(L.mcb200_workspace_counters(ws, cnt))
        assert cnt[0] + cnt[1] + cnt[2] == nq and cnt[1] > 0 and cnt[2] > 0
        check_tops()
    finally:
        g1.db.set_target_taxa(None)
    L.mcb200_workspace_destroy(ws)


def test_cpp_shim_runs_the_reference_call_sequence(g1, tmp_path):
    """metacache_b200/host/shim_query.cpp = database_query.hpp:87-124 written against the C++
    shims (gpu_hashmap / query_batch mirrors); its output must equal the reference's"""
    import subprocess
    from metacache_b200 import dbformat
    from oracle import refio
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "metacache_b200", "host", "shim_query")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    dbp = str(tmp_path / "g1.cache0")
    dbformat.write_cache(dbp, dbformat.CachePart(g1.keys, g1.sizes, g1.values))
    rp = str(tmp_path / "reads.txt")
    norm = lambda x: x if len(x) else b"-"
    refio.write_reads_txt(rp, [(norm(a), norm(b)) if len(b) else norm(a) for a, b in g1.reads])
    out = subprocess.run([exe, dbp, rp, "2", "1"], check=True, capture_output=True, text=True).stdout.splitlines()
    out = [ln for ln in out if not ln.startswith("#")]
    exp = g1.expected("c2_")
    assert len(out) == len(g1.reads)
    for i, line in enumerate(out):
        qid, tops, nall = line.split("\t")
        assert int(qid) == i
        got = [tuple(int(x) for x in t.split(":")) for t in tops.split(",")] if tops else []
        assert got == exp.top[i], i
        assert int(nall) == len(exp.allhits[i]), i


def test_wide_location_layout_gives_identical_results(g1, monkeypatch):
    """tables are stored with 32-bit packed locations when ids fit; the 64-bit layout (forced
    here) must behave identically"""
    from metacache_b200.database import Database, query_reads
    monkeypatch.setenv("MCB200_WIDE_LOCATIONS", "1")
    db = Database(0, 1)
    db.load_part_arrays(0, g1.keys, g1.sizes, g1.values)
    monkeypatch.delenv("MCB200_WIDE_LOCATIONS")
    exp = g1.expected("c2_")
    res = query_reads(db, g1.reads, _sk(g1), copy_all_hits=True)
    for i, (allh, top) in enumerate(res):
        assert np.array_equal(allh, exp.allhits[i]), i
        assert top == exp.top[i], i
    res = query_reads(db, g1.reads, _sk(g1), copy_all_hits=False)
    assert [r[1] for r in res] == exp.top
    k, s, v = db.export_part(0)
    o, ro = np.argsort(k), np.argsort(g1.keys)
    assert np.array_equal(k[o], g1.keys[ro]) and np.array_equal(s[o], g1.sizes[ro])
    db.close()


def test_export_round_trip_packed_layout(g1):
    k, s, v = g1.db.export_part(0)
    o, ro = np.argsort(k), np.argsort(g1.keys)
    assert np.array_equal(k[o], g1.keys[ro]) and np.array_equal(s[o], g1.sizes[ro])
    offs = np.zeros(len(s) + 1, np.int64); np.cumsum(s, out=offs[1:])
    roffs = np.zeros(len(g1.sizes) + 1, np.int64); np.cumsum(g1.sizes, out=roffs[1:])
    for a, b in zip(o, ro):
        assert np.array_equal(v[offs[a]:offs[a + 1]], g1.values[roffs[b]:roffs[b + 1]])


def test_classify_on_device_matches_oracle_with_synthetic_lineages(g1):
    """classify() with a made-up taxonomy over the golden targets: exercises LCA at several ranks,
    the hit-difference threshold and the highest-rank cut-off"""
    from metacache_b200.database import query_reads
    from oracle import mc_oracle as O
    nt = len(g1.target_windows)
    lin = np.zeros((nt, 21), np.uint32)
    for t in range(nt):
        lin[t, 0] = 1 + t                        # sequence level: the target itself
        fam = {10: 0, 11: 0, 12: 3, 13: 3, 14: 7, 15: 7}.get(t, t)
        lin[t, 4] = 100 + fam                    # species: G0/G0m1/G0m2 share one, ...
        if t % 3:
            lin[t, 6] = 200 + fam // 4           # genus missing for every third target
        lin[t, 16] = 300 + (fam % 2)             # phylum
        lin[t, 19] = 400                         # domain
    g1.db.copy_target_lineages_to_gpus(lin)
    tab = O.Table(g1.keys, g1.sizes, g1.values)
    for frac, highest, maxc in ((1.0, 19, 2), (0.3, 19, 4), (1.0, 6, 4), (0.0, 16, 5)):
        res = query_reads(g1.db, g1.reads, _sk(g1), max_candidates=maxc, copy_all_hits=False, classify=True,
                          hits_diff_fraction=frac, highest_rank=highest)
        n_cls = 0
        for i, (a, b) in enumerate(g1.reads):
            _, top = O.query(tab, a, b, maxc=maxc)
            want = O.classify(top, lin, hits_min=5, hits_diff_fraction=frac, lowest=0, highest=highest)
            got = res[i][2]
            if want[0] == 0:
                assert got[0] == 0, (i, frac, highest)
            else:
                assert got == want, (i, frac, highest)
                n_cls += 1
        assert n_cls > 300


@pytest.mark.parametrize("k,s,w,stride", [
    (16, 16, 127, 112),      # default geometry: thread-per-window kernel
    (12, 8, 60, 49),         # short k-mers, short sketches (masked k-mer roll, scalar row store)
    (16, 16, 256, 241),      # longest window the thread-per-window kernel stages
    (9, 16, 300, 292),       # longer windows: warp-per-window kernel
    (16, 32, 127, 112),      # sketches of 32 features: warp-per-window kernel
    (16, 16, 40, 10),        # heavily overlapping windows, windows of <= 32 k-mers only
])
def test_sketch_geometries_against_oracle(g1, k, s, w, stride):
    """window sketches for several (k, s, w, stride) vs the oracle's for_each_sketch restatement:
    random reads with N runs, tandem repeats (duplicate k-mers inside a window), homopolymers,
    reads shorter than k / w and lower case"""
    from metacache_b200.database import SketchingOpt, query_reads
    from oracle import mc_oracle as O
    rng = np.random.default_rng(k * 1000 + w)
    reads = []
    for i in range(400):
        n = int(rng.choice([1, 5, k - 1, k, k + 1, w - 1, w, w + 1, 150, 151, 333, 1000, 2500]))
        r = rng.choice(list(b"ACGT"), n).astype(np.uint8)
        kind = i % 8
        if kind == 1 and n:
            r[rng.integers(0, n, max(1, n // 50))] = ord("N")
        elif kind == 2 and n > 40:
            a = int(rng.integers(0, n - 20)); r[a:a + int(rng.integers(1, 40))] = ord("N")
        elif kind == 3 and n:
            unit = rng.choice(list(b"ACGT"), int(rng.integers(1, 7))).astype(np.uint8)
            r = np.resize(unit, n)
        elif kind == 4 and n:
            r[:] = ord("A")
        elif kind == 5:
            r = np.frombuffer(bytes(r).lower(), np.uint8).copy()
        reads.append(r.tobytes())
    res = query_reads(g1.db, reads, SketchingOpt(k, s, w, stride), with_sketches=True, copy_all_hits=False)
    for i, (seq, item) in enumerate(zip(reads, res)):
        want = O.sketch_sequence(seq, k, s, w, stride)
        got = item[2]
        assert len(got) == len(want), (i, len(seq))
        for j, (x, y) in enumerate(zip(got, want)):
            y = np.zeros(0, np.uint32) if y is None else y
            assert np.array_equal(x, y), (i, j, len(seq))


def _tops_of(top_rows):
    return [tuple(int(x) for x in row) for row in top_rows if row[1] > 0]


@pytest.mark.parametrize("threads", [1, 3])
def test_query_file_matches_query_reads(g1, tmp_path, threads):
    """file -> reader threads -> batch slots -> top hits == add_paired_read per read;
    FASTA (wrapped lines), FASTQ and two paired files"""
    from metacache_b200.database import query_reads
    from metacache_b200.reader import query_file
    sk = _sk(g1)
    # single-end reads; FASTA cannot express an empty first line distinctly, so keep reads with characters
    single = [a for a, b in g1.reads if len(a) > 0 and len(b) == 0]
    want = [r[1] for r in query_reads(g1.db, single, sk, copy_all_hits=False)]
    fa = tmp_path / "single.fa"
    with open(fa, "wb") as f:
        for i, s in enumerate(single):
            f.write(b">read_%d some description\n" % i)
            for p in range(0, len(s), 70):
                f.write(s[p:p + 70] + b"\n")
    top, heads = query_file(g1.db, str(fa), sketching=sk, threads=threads, batch_queries=37, keep_headers=True)
    assert len(top) == len(single)
    assert [h.decode() for h in heads] == ["read_%d some description" % i for i in range(len(single))]
    assert [_tops_of(t) for t in top] == want
    fq = tmp_path / "single.fq"
    with open(fq, "wb") as f:
        for i, s in enumerate(single):
            f.write(b"@read_%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
    top, _ = query_file(g1.db, str(fq), sketching=sk, threads=threads, batch_queries=1000)
    assert [_tops_of(t) for t in top] == want
    # pairs
    pairs = [(a, b) for a, b in g1.reads if len(a) > 0 and len(b) > 0]
    if pairs:
        want_p = [r[1] for r in query_reads(g1.db, pairs, sk, copy_all_hits=False)]
        f1, f2 = tmp_path / "p.1.fa", tmp_path / "p.2.fa"
        with open(f1, "wb") as a, open(f2, "wb") as b:
            for i, (x, y) in enumerate(pairs):
                a.write(b">p%d/1\n" % i + x + b"\n")
                b.write(b">p%d/2\n" % i + y + b"\n")
        top, _ = query_file(g1.db, str(f1), str(f2), sketching=sk, threads=threads, batch_queries=50)
        assert [_tops_of(t) for t in top] == want_p
        # the same pairs interleaved in one file (-pairseq)
        il = tmp_path / "p.il.fa"
        with open(il, "wb") as f:
            for i, (x, y) in enumerate(pairs):
                f.write(b">p%d/1\n" % i + x + b"\n>p%d/2\n" % i + y + b"\n")
        top, _ = query_file(g1.db, str(il), str(il), sketching=sk, threads=threads, batch_queries=50)
        assert [_tops_of(t) for t in top] == want_p


def test_cpp_query_files_driver(g1, tmp_path):
    """metacache_b200/host/shim_query_file.cpp: mcb200::query_files (C++ reader + worker threads over
    the C ABI, the query_batched replacement) on a FASTA file == add_paired_read per read"""
    import subprocess
    from metacache_b200 import dbformat
    from metacache_b200.database import query_reads
    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "metacache_b200", "host")
    exe = os.path.join(host, "shim_query_file")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", host])
    dbp = str(tmp_path / "g1.cache0")
    dbformat.write_cache(dbp, dbformat.CachePart(g1.keys, g1.sizes, g1.values))
    single = [a for a, b in g1.reads if len(a) > 0 and len(b) == 0]
    want = [r[1] for r in query_reads(g1.db, single, _sk(g1), copy_all_hits=False)]
    fa = str(tmp_path / "single.fa")
    with open(fa, "wb") as f:
        for i, s in enumerate(single):
            f.write(b">q%d\n" % i)
            for p in range(0, len(s), 80):
                f.write(s[p:p + 80] + b"\n")
    for threads in ("1", "5"):
        out = subprocess.run([exe, dbp, fa, "-", threads, "2"], check=True, capture_output=True, text=True).stdout.splitlines()
        assert len(out) == len(single)
        for i, line in enumerate(out):
            head, tops = line.split("\t")
            assert head == "q%d" % i
            got = [tuple(int(x) for x in t.split(":")) for t in tops.split(",")] if tops else []
            assert got == want[i], i


@pytest.mark.parametrize("gi", range(5))
def test_other_geometries_match_reference(gi):
    """databases built and queried by the reference with non-default (k, s, w, stride)
    (tests/golden/g3.npz): sketches, all-hits and top candidates of the CUDA path"""
    from metacache_b200.database import Database, SketchingOpt, query_reads
    from tests.golden_util import G3
    g3 = G3()
    k, s, w, stride = g3.geometries[gi]
    exp = g3.expected(gi)
    db = Database(0, 1)
    db.load_part_arrays(0, *g3.part(gi))
    # the rules of make_candidate_generation_rules use the DATABASE's window stride
    from metacache_b200 import database as D
    res = []
    qb = D.QueryBatch(db, 4096, 1 << 22, 2, True, 1)
    hd = qb.host_data(0)
    sk = SketchingOpt(k, s, w, stride)
    for a, b in g3.reads:
        rules = D.make_candidate_generation_rules(len(a), len(b), 0, stride, 2)
        assert qb.add_paired_read(0, a, b, sk, rules)
    db.query_gpu_async(qb, 0, sk)
    hd.wait_for_results()
    for i, (a, b) in enumerate(g3.reads):
        kept = [x for x, keep in zip(hd.sketches(i), _window_keeps(a, b, k, w, stride)) if keep]
        assert len(kept) == len(exp.sketches[i]), i
        for x, y in zip(kept, exp.sketches[i]):
            assert np.array_equal(x, y), i
        assert np.array_equal(hd.allhits(i), exp.allhits[i]), i
        top = [c.as_tuple() for c in hd.top_candidates(i) if c.hits > 0]
        assert top == exp.top[i], i
    qb.close()
    db.close()


def test_host_packed_bases_equal_device_encoding(g1):
    """mcb200_pack_bases (host, csrc/pack.cpp) + mcb200_query_packed_device == encode_kernel +
    mcb200_query_device on the edge-case reads of G1 (IUPAC, N runs, lower case, U, empty mates):
    same sketches, same candidates."""
    import ctypes as C
    import torch
    from metacache_b200 import _lib
    from metacache_b200._lib import DevQueries, Sketching
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    seqs, seq_query = [], []
    for qi, (a, b) in enumerate(g1.reads):
        mates = [m for m in (a, b) if len(m)] or [b""]
        for m in mates:
            seqs.append(m)
            seq_query.append(qi)
    lens = np.array([len(s) for s in seqs], np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)])
    flat = np.frombuffer(b"".join(seqs), np.uint8)
    nq, ns, nb = len(g1.reads), len(seqs), int(offs[-1])
    units = (nb + 31) // 32
    codes = np.zeros(2 * (units + 1) + 16, np.uint32)
    amb = np.zeros(units + 1 + 16, np.uint32)
    pos = 0
    for s in seqs:                                             # read by read: every bit offset occurs
        buf = np.frombuffer(s, np.uint8)
        assert L.mcb200_pack_bases(buf.ctypes.data if len(buf) else None, len(buf), pos, codes.ctypes.data, amb.ctypes.data) >= 0
        pos += len(buf)
    if nb & 31:
        amb[units - 1] |= np.uint32(0xFFFFFFFF >> (nb & 31))
    pad = np.zeros(64, np.uint8)
    d_flat = torch.from_numpy(np.concatenate([flat, pad])).to(dev)
    d_codes, d_amb = torch.from_numpy(codes.view(np.int32)).to(dev), torch.from_numpy(amb.view(np.int32)).to(dev)
    d_off = torch.from_numpy(offs.astype(np.uint32).view(np.int32)).to(dev)
    d_sq = torch.from_numpy(np.array(seq_query, np.int32)).to(dev)
    mw = np.array([2 + max(len(a) + len(b), 0) // g1.stride for a, b in g1.reads], np.int32)
    d_mw = torch.from_numpy(mw).to(dev)
    sk = Sketching(g1.k, g1.s, g1.w, g1.stride)
    out = []
    for packed in (False, True):
        ws = _lib.check_ptr(L.mcb200_workspace_create(g1.db._h, nq, ns, nb + 64, 2, 0))
        q = DevQueries(None if packed else d_flat.data_ptr(), d_off.data_ptr(), d_sq.data_ptr(), d_mw.data_ptr(), ns, nq, nb)
        top = torch.empty((nq, 2, 4), dtype=torch.int32, device=dev)
        for attempt in range(6):
            if packed:
                _lib.check(L.mcb200_query_packed_device(ws, C.byref(q), d_codes.data_ptr(), d_amb.data_ptr(),
                                                        C.byref(sk), top.data_ptr(), None))
            else:
                _lib.check(L.mcb200_query_device(ws, C.byref(q), C.byref(sk), top.data_ptr(), None))
            rc = L.mcb200_workspace_check(ws)
            if rc != _lib.EAGAIN:
                _lib.check(rc)
                break
        nw = L.mcb200_workspace_num_windows(ws)
        from metacache_b200.distributed import _as_tensor
        feats = _as_tensor(L.mcb200_workspace_sketches(ws), nw * g1.s, dev).cpu().numpy().copy()
        out.append((top.cpu().numpy().copy(), feats))
        L.mcb200_workspace_destroy(ws)
    assert np.array_equal(out[0][1], out[1][1]), "sketches differ between device-encoded and host-packed bases"
    assert np.array_equal(out[0][0], out[1][0])
    assert (out[0][0][:, 0, 1] > 0).sum() > 100
