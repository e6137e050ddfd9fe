"""Parity at benchmark scale: seeded synthetic workloads (BASELINE C2 recipe) checked through
size-independent properties on every read, and bit-exactly against the oracle on a sample."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SK = dict(kmerlen=16, sketchlen=16, winlen=127, winstride=112)
RL, MAXC = 150, 2


def _build(n_targets, target_len, device, want_windows=False):
    import torch
    from metacache_b200 import _lib, synth
    from metacache_b200._lib import Sketching
    from metacache_b200.database import Database
    bases, off = synth.make_targets(n_targets, target_len, 10, synth.SEED_DB, device=device)
    db = Database(device.index, 1)
    sk = Sketching(**SK)
    wins = np.zeros(n_targets, np.uint32)
    _lib.check(_lib.lib().mcb200_db_build_part_from_targets(db._h, 0, bases.data_ptr(), off.data_ptr(), n_targets, 0,
                                                            C.byref(sk), 254, 0.0, wins.ctypes.data))
    torch.cuda.synchronize(device)
    if want_windows:
        return db, bases, wins
    return db, bases


def _query_device(db, reads, device):
    """whole batch through mcb200_query_device -> [nq, MAXC, 4] uint32 numpy"""
    import torch
    from metacache_b200 import _lib
    from metacache_b200._lib import DevQueries, Sketching
    L = _lib.lib()
    nq = reads.shape[0]
    flat = reads.reshape(-1)
    seq_off = (torch.arange(nq + 1, dtype=torch.int64, device=device) * RL).to(torch.int32)
    seq_qry = torch.arange(nq, dtype=torch.int32, device=device)
    max_win = torch.full((nq,), 2 + RL // 112, dtype=torch.int32, device=device)
    ws = _lib.check_ptr(L.mcb200_workspace_create(db._h, nq, nq, nq * RL + 64, MAXC, 0))
    q = DevQueries(flat.data_ptr(), seq_off.data_ptr(), seq_qry.data_ptr(), max_win.data_ptr(), nq, nq, nq * RL)
    sk = Sketching(**SK)
    top = torch.empty((nq, MAXC, 4), dtype=torch.int32, device=device)
    _lib.query_device_checked(ws, q, sk, top.data_ptr())
    out = top.cpu().numpy().view(np.uint32)
    L.mcb200_workspace_destroy(ws)
    return out


def _query_batches(db, reads_np, slot_reads):
    """the same reads through the host batch API in slots of slot_reads"""
    from metacache_b200 import _lib
    from metacache_b200._lib import Sketching
    L = _lib.lib()
    nq = reads_np.shape[0]
    nslots = (nq + slot_reads - 1) // slot_reads
    qb = _lib.check_ptr(L.mcb200_batch_create(db._h, slot_reads, slot_reads * RL + 64, MAXC, 0, nslots))
    offs = np.arange(slot_reads + 1, dtype=np.uint64) * RL
    sk = Sketching(**SK)
    flat = reads_np.reshape(-1)
    for s in range(nslots):
        n = min(slot_reads, nq - s * slot_reads)
        chunk = np.ascontiguousarray(flat[s * slot_reads * RL:(s * slot_reads + n) * RL])
        assert _lib.check(L.mcb200_batch_add_reads(qb, s, chunk.ctypes.data, offs.ctypes.data, n, 0, 0, 112)) == n
        _lib.check(L.mcb200_batch_submit(qb, s, C.byref(sk)))
    out = np.zeros((nq, MAXC, 4), np.uint32)
    for s in range(nslots):
        _lib.check(L.mcb200_batch_wait(qb, s))
        n = L.mcb200_batch_num_queries(qb, s)
        p = L.mcb200_batch_top_candidates(qb, s, 0)
        out[s * slot_reads:s * slot_reads + n] = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n, MAXC, 4))
    L.mcb200_batch_destroy(qb)
    return out


def _check_properties(top, n_targets):
    hits = top[:, :, 1].astype(np.int64)
    tgt, beg, end = top[:, :, 0], top[:, :, 2], top[:, :, 3]
    used = hits > 0
    assert np.all(hits[:, 0] >= hits[:, 1])                          # sorted by hits, best first
    assert np.all(~used[:, 1] | used[:, 0])                          # no hole in the list
    assert np.all(tgt[used] < n_targets) and np.all(tgt[~used] == 0xFFFFFFFF)
    assert np.all(beg[used] <= end[used]) and np.all(end[used] - beg[used] < 3)    # W = 2 + 150//112 = 3
    assert np.all(hits <= 32 * 3)                                    # <= features x windows in range
    both = used[:, 0] & used[:, 1]
    assert np.all(tgt[both, 0] != tgt[both, 1])                      # distinct targets
    tie = both & (hits[:, 0] == hits[:, 1])
    assert np.all(tgt[tie, 0] < tgt[tie, 1])                         # stable order: lower target first on ties


def test_medium_scale_properties_and_oracle_sample():
    import torch
    from metacache_b200 import synth
    from oracle import mc_oracle as O
    device = torch.device("cuda", 0)
    NT, TL, NQ = 5000, 100_000, 2_000_000
    db, bases = _build(NT, TL, device)
    reads = synth.make_reads_150(NQ, bases, NT, TL, RL, device=device)
    top = _query_device(db, reads, device)
    _check_properties(top, NT)
    # reads sampled from the database find their source family (10 consecutive targets) first
    top2 = _query_device(db, reads, device)
    assert np.array_equal(top, top2)                                 # idempotent / deterministic
    # host batch API in ragged slots == device API on the whole batch
    reads_np = reads.cpu().numpy()
    assert np.array_equal(_query_batches(db, reads_np[:300_001], 77_777), top[:300_001])
    # bit-exact against the oracle on a seeded sample
    keys, sizes, values = db.export_part(0)
    tab = O.Table(keys, sizes, values)
    rng = np.random.default_rng(11)
    for i in rng.choice(NQ, 4000, replace=False):
        _, want = O.query(tab, reads_np[i].tobytes(), b"")
        got = [tuple(int(x) for x in row) for row in top[i] if row[1] > 0]
        assert got == want, int(i)
    mapped = (top[:, 0, 1] >= 5).mean()
    assert mapped > 0.85                                             # 90 % of R150 comes from the database
    db.close()


def _reference_tops(db, wins, reads_np, tmp, first=0):
    """top candidates of `reads_np` ([n, RL] uint8) from the REFERENCE's own hot path: the part is
    exported to the reference's .meta/.cache0 format and queried by oracle/_ref/mc_ref_harness
    (unmodified reference objects, database::query_host) on all host threads"""
    from metacache_b200 import dbformat
    from oracle import refio
    assert os.path.exists(refio.HARNESS), "oracle/_ref/mc_ref_harness missing: run build() where /root/reference exists"
    base = os.path.join(tmp, "db")
    keys, sizes, values = db.export_part(0)
    dbformat.write_cache(base + ".cache0", dbformat.CachePart(keys, sizes, values))
    dbformat.write_meta(base + ".meta", dbformat.synthetic_meta(wins, **SK))
    del keys, sizes, values
    rt = os.path.join(tmp, "reads.txt")
    lines = np.full((reads_np.shape[0], RL + 1), ord("\n"), np.uint8)
    lines[:, :RL] = reads_np
    lines.tofile(rt)
    tops = os.path.join(tmp, "tops.bin")
    refio.run_harness(base, rt, "-", threads=os.cpu_count() or 1, repeat=1, sketches=0, allhits=0, maxcand=MAXC, tops=tops)
    ref = np.fromfile(tops, dtype="<u4").reshape(-1, MAXC, 4)
    for f in (base + ".cache0", base + ".meta", rt, tops):
        os.unlink(f)
    return ref


def _scratch_dir():
    d = "/dev/shm" if os.path.isdir("/dev/shm") else None
    return tempfile.mkdtemp(prefix="mcb200_scale_", dir=d)


def test_full_scale_c2_properties_and_reference_sample():
    """BASELINE config C2 at full size: 10 M reads vs 50 k targets (714 M locations).  Properties on
    every read, and the first 400 k reads bit-exactly against the reference's own hot path on the very
    same full-size database (the headline configuration was only property-checked)."""
    import torch
    from metacache_b200 import synth
    device = torch.device("cuda", 0)
    NT, TL, NQ = 50_000, 100_000, 10_000_000
    db, bases, wins = _build(NT, TL, device, want_windows=True)
    assert db.value_count(0) == 714_400_000
    reads = synth.make_reads_150(NQ, bases, NT, TL, RL, device=device)
    del bases
    top = _query_device(db, reads, device)
    _check_properties(top, NT)
    NREF = 400_000
    tmp = _scratch_dir()
    try:
        ref = _reference_tops(db, wins, reads[:NREF].cpu().numpy(), tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    bad = np.flatnonzero((ref != top[:NREF]).any(axis=(1, 2)))
    assert len(bad) == 0, (len(bad), int(bad[0]), ref[bad[0]].tolist(), top[bad[0]].tolist())
    assert (ref[:, 0, 1] >= 5).mean() > 0.85
    # checksum of checksums: a second pass over the two halves in swapped order gives the same rows
    half = NQ // 2
    swapped = torch.cat([reads[half:], reads[:half]])
    top_s = _query_device(db, swapped, device)
    assert np.array_equal(top_s[:NQ - half], top[half:]) and np.array_equal(top_s[NQ - half:], top[:half])
    # source family of a sampled read = its best hit's family for (almost) every mapped read
    h = (top[:, 0, 1] >= 5)
    assert 0.85 < h.mean() < 0.95
    db.close()


def test_long_reads_c3_properties_and_oracle_sample():
    """BASELINE config C3 (long reads, 200-19000 bp) at reduced size: reads of every length class go
    through the fused kernel (<= 8 windows per range), the CTA kernel (longer) or its global-memory
    variant; a seeded sample incl. the longest reads is compared bit-exactly with the oracle."""
    import torch
    from metacache_b200 import _lib, synth
    from metacache_b200._lib import DevQueries, Sketching
    from oracle import mc_oracle as O
    device = torch.device("cuda", 0)
    NT, TL, NQ = 2000, 100_000, 60_000
    db, bases = _build(NT, TL, device)
    flat, offs = synth.make_long_reads(NQ, bases, NT, TL, device=device)
    lens = (offs[1:] - offs[:-1])
    L = _lib.lib()
    n_bases = int(offs[-1].item())
    seq_off = offs.to(torch.int32)
    seq_qry = torch.arange(NQ, dtype=torch.int32, device=device)
    max_win = (2 + lens // 112).to(torch.int32)
    ws = _lib.check_ptr(L.mcb200_workspace_create(db._h, NQ, NQ, n_bases + 64, MAXC, 0))
    q = DevQueries(flat.data_ptr(), seq_off.data_ptr(), seq_qry.data_ptr(), max_win.data_ptr(), NQ, NQ, n_bases)
    sk = Sketching(**SK)
    top_d = torch.empty((NQ, MAXC, 4), dtype=torch.int32, device=device)
    _lib.query_device_checked(ws, q, sk, top_d.data_ptr())
    cnt = (C.c_uint64 * 8)()
    _lib.check(L.mcb200_workspace_counters(ws, cnt))
    top = top_d.cpu().numpy().view(np.uint32)
    L.mcb200_workspace_destroy(ws)

    hits = top[:, :, 1].astype(np.int64)
    used = hits > 0
    lens_np = lens.cpu().numpy()
    W = 2 + lens_np // 112
    assert np.all(hits[:, 0] >= hits[:, 1]) and np.all(~used[:, 1] | used[:, 0])
    assert np.all(top[:, :, 0][used] < NT)
    span = (top[:, :, 3].astype(np.int64) - top[:, :, 2].astype(np.int64))
    assert np.all(span[used] >= 0) and np.all((span < W[:, None])[used])
    both = used[:, 0] & used[:, 1]
    assert np.all(top[both, 0, 0] != top[both, 1, 0])
    assert (hits[:, 0] >= 5).mean() > 0.95                       # every long read comes from the database

    keys, sizes, values = db.export_part(0)
    tab = O.Table(keys, sizes, values)
    flat_np, offs_np = flat.cpu().numpy(), offs.cpu().numpy()
    rng = np.random.default_rng(12)
    sample = set(int(i) for i in rng.choice(NQ, 1500, replace=False)) | set(int(i) for i in np.argsort(-lens_np)[:40])
    for i in sorted(sample):
        _, want = O.query(tab, flat_np[offs_np[i]:offs_np[i + 1]].tobytes(), b"")
        got = [tuple(int(x) for x in row) for row in top[i] if row[1] > 0]
        assert got == want, (i, int(lens_np[i]))
    db.close()


def test_device_builder_matches_reference_build_on_1000_dbs_targets():
    """The database both bench arms query is made by mcb200_db_build_part_from_targets.  Here the same
    DB-S recipe, 1 000 targets x 100 kbp (14.3 M locations), is ALSO built by the reference's own
    `metacache build` from a FASTA file of the same sequences: keys, bucket sizes and every bucket's
    (tgt,win)-ordered location list must be identical (equivalence had only been checked on
    a 16-target FASTA)."""
    import torch
    from metacache_b200 import dbformat, synth
    from oracle import refio
    assert os.path.exists(refio.METACACHE), "oracle/_ref/metacache missing: run build() where /root/reference exists"
    device = torch.device("cuda", 0)
    NT, TL = 1000, 100_000
    db, bases, wins = _build(NT, TL, device, want_windows=True)
    keys, sizes, values = db.export_part(0)
    tmp = _scratch_dir()
    try:
        fa = os.path.join(tmp, "t.fa")
        seqs = bases.cpu().numpy().reshape(NT, TL)
        with open(fa, "wb") as f:
            for i in range(NT):
                f.write(b">t%d\n" % i)
                f.write(seqs[i].tobytes())
                f.write(b"\n")
        subprocess.check_call([refio.METACACHE, "build", os.path.join(tmp, "ref"), fa, "-parts", "1", "-silent"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ref = dbformat.read_cache(os.path.join(tmp, "ref.cache0"))
        meta = dbformat.read_meta(os.path.join(tmp, "ref.meta"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    assert meta.target_count == NT and np.array_equal(np.asarray(meta.target_windows(), np.uint32), wins)
    assert len(keys) == len(ref.keys) and len(values) == len(ref.values) and len(values) > 14_000_000
    o, ro = np.argsort(keys), np.argsort(ref.keys)
    assert np.array_equal(keys[o], ref.keys[ro]) and np.array_equal(sizes[o], ref.sizes[ro])
    # location lists in key order: gather both value arrays bucket by bucket
    def by_key(sz, vals, order):
        offs = np.zeros(len(sz) + 1, np.int64)
        np.cumsum(sz, out=offs[1:])
        lens = sz[order].astype(np.int64)
        start = np.repeat(offs[:-1][order], lens)
        within = np.arange(lens.sum(), dtype=np.int64) - np.repeat(np.cumsum(lens) - lens, lens)
        return vals[start + within]
    assert np.array_equal(by_key(sizes, values, o), by_key(ref.sizes, ref.values, ro))
    db.close()
