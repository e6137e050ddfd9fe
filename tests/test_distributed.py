"""Host-side logic of the multi-GPU paths on the CPU.  The compute steps are injected
(tests/shard_numpy_backend.py: numpy + the oracle); what runs for real is
metacache_b200.distributed.FeatureShardedQuery - per-owner split sizes, segment bookkeeping, the
chunk pipeline and the collectives - over gloo with 2 processes, and over ThreadComm with 3 ranks in
one process.  Expected values are the REFERENCE's per-part outputs (golden g2: a 2-part database
built and queried by the reference) merged in part order (docs/partitioning.md:116-142)."""
import os
import socket
import threading

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.golden_util import G1, G2

MAXC, S = 2, 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sketch_reads(reads):
    """-> feats uint32 [nwin, S] (0xFFFFFFFF padded), qwo [n + 1], max_win int32 [n]"""
    from oracle import mc_oracle as O
    rows, qwo, mw = [], [0], []
    for a, b in reads:
        for seq in (a, b):
            for x in O.sketch_sequence(seq):
                if x is None:
                    continue
                r = np.full(S, 0xFFFFFFFF, np.uint32)
                r[:len(x)] = x
                rows.append(r)
        qwo.append(len(rows))
        mw.append(O.max_windows_in_range(len(a), len(b)))
    feats = np.stack(rows) if rows else np.zeros((0, S), np.uint32)
    return feats, np.array(qwo, np.int64), torch.tensor(mw, dtype=torch.int32)


def _run_rank(comm, rank, world, reads, chunk):
    from metacache_b200.distributed import FeatureShardedQuery
    from tests.shard_numpy_backend import NumpyBackend
    g2 = G2()
    per = (len(reads) + world - 1) // world
    mine = reads[rank * per:(rank + 1) * per]           # the last rank may hold fewer reads (or none)
    feats, qwo, max_win = _sketch_reads(mine)
    backend = NumpyBackend(g2.parts, rank, world, MAXC)
    fq = FeatureShardedQuery(backend, comm, S, MAXC, chunk_queries=chunk, n_slots=3, streams=None)
    top = torch.zeros((len(mine), MAXC, 4), dtype=torch.int32)

    def sketches(q0, q1):
        return feats, qwo[q0:q1 + 1], max_win[q0:q1]

    fq.step(len(mine), sketches, top, feat_cap=max(len(feats) * S, 1))
    out = []
    for row in top.numpy().view(np.uint32):
        out.append([tuple(int(x) for x in c) for c in row if c[1] > 0])
    return out, dict(fq.stats)


def _expected(n):
    from oracle import mc_oracle as O
    g2 = G2()
    e0, e1 = g2.expected(0), g2.expected(1)
    return [O.merge_tops([e0.top[i], e1.top[i]], MAXC) for i in range(n)]


def _worker(rank, world, port, nreads, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from metacache_b200.distributed import TorchComm
    got, stats = _run_rank(TorchComm(), rank, world, G1().reads[:nreads], chunk=23)
    torch.save((got, stats), os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_over_gloo_match_the_reference_per_part_merge(tmp_path):
    world, nreads = 2, 150
    mp.spawn(_worker, args=(world, _free_port(), nreads, str(tmp_path)), nprocs=world, join=True)
    got, feats_sent = [], 0
    for r in range(world):
        g, st = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        got += g
        feats_sent += st["features_sent"]
        assert st["chunks"] == 4                         # 75 reads in chunks of 23
    want = _expected(nreads)
    assert len(got) == nreads and feats_sent > 1000
    for i in range(nreads):
        assert got[i] == want[i], i
    assert sum(1 for t in want if t) > 50


@pytest.mark.parametrize("world,nreads,chunk", [(3, 100, 16), (2, 61, 1000), (4, 3, 2)])
def test_ranks_as_threads_match_the_reference_per_part_merge(world, nreads, chunk):
    """ThreadComm: ragged slices (the last ranks hold fewer reads or none) and a single-chunk step"""
    from metacache_b200.distributed import ThreadComm
    shared = ThreadComm.Shared(world)
    reads = G1().reads[:nreads]
    res, errs = [None] * world, []

    def run(r):
        try:
            res[r] = _run_rank(ThreadComm(shared, r), r, world, reads, chunk)[0]
        except BaseException as ex:                      # noqa: BLE001 - re-raised below
            errs.append(ex)
            shared.barrier.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]
    got = [x for r in res for x in r]
    want = _expected(nreads)
    assert len(got) == nreads
    for i in range(nreads):
        assert got[i] == want[i], i


def test_owner_function_matches_the_library():
    """shard_of restated in numpy (tests/shard_numpy_backend.py) spreads features evenly; the GPU
    tests compare it with the device's routing"""
    from tests.shard_numpy_backend import shard_of_np
    keys = G2().parts[0][0]
    for n in (2, 3, 8):
        o = shard_of_np(keys, n)
        assert o.min() == 0 and o.max() == n - 1
        c = np.bincount(o, minlength=n)
        assert c.max() < 1.2 * c.min()
