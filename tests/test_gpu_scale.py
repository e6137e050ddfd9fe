"""Parity at benchmark scale: seeded synthetic workloads (BASELINE C2 recipe) checked through
size-independent properties on every read, and bit-exactly against the oracle on a sample."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SK = dict(kmerlen=16, sketchlen=16, winlen=127, winstride=112)
RL, MAXC = 150, 2


def _build(n_targets, target_len, device):
    import torch
    from metacache_b200 import _lib, synth
    from metacache_b200._lib import Sketching
    from metacache_b200.database import Database
    bases, off = synth.make_targets(n_targets, target_len, 10, synth.SEED_DB, device=device)
    db = Database(device.index, 1)
    sk = Sketching(**SK)
    _lib.check(_lib.lib().mcb200_db_build_part_from_targets(db._h, 0, bases.data_ptr(), off.data_ptr(), n_targets, 0,
                                                            C.byref(sk), 254, 0.0, None))
    torch.cuda.synchronize(device)
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
    _lib.check(L.mcb200_query_device(ws, C.byref(q), C.byref(sk), top.data_ptr(), None))
    torch.cuda.synchronize(device)
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


def test_full_scale_c2_properties():
    """BASELINE config C2 at full size: 10 M reads vs 50 k targets (714 M locations)"""
    import torch
    from metacache_b200 import synth
    device = torch.device("cuda", 0)
    NT, TL, NQ = 50_000, 100_000, 10_000_000
    db, bases = _build(NT, TL, device)
    assert db.value_count(0) == 714_400_000
    reads = synth.make_reads_150(NQ, bases, NT, TL, RL, device=device)
    del bases
    top = _query_device(db, reads, device)
    _check_properties(top, NT)
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
    _lib.check(L.mcb200_query_device(ws, C.byref(q), C.byref(sk), top_d.data_ptr(), None))
    torch.cuda.synchronize(device)
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
