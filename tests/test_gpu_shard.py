"""Feature-space sharding on the device (csrc/kernels_shard.cu, metacache_b200/distributed.py).
One GPU is enough: the shards are the ranks of a ThreadComm (N threads of this process, each with its
own shard table on cuda:0), so route / probe / gather / reduce and the whole exchange choreography run
exactly as over NCCL.  Expected values: the REFERENCE's per-part outputs for the 2-part database of
golden g2, merged in part order (docs/partitioning.md:116-142)."""
import ctypes as C
import threading

import numpy as np
import pytest

from tests.golden_util import G1, G2

pytestmark = pytest.mark.gpu
MAXC = 2


def _expected(n):
    from oracle import mc_oracle as O
    g2 = G2()
    e0, e1 = g2.expected(0), g2.expected(1)
    return [O.merge_tops([e0.top[i], e1.top[i]], MAXC) for i in range(n)]


def _run_threads(world, fn):
    from metacache_b200.distributed import ThreadComm
    shared = ThreadComm.Shared(world)
    res, errs = [None] * world, []

    def run(r):
        try:
            res[r] = fn(ThreadComm(shared, r), r)
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
    return res


def _rank(comm, rank, world, reads, chunk, g1, parts, n_targets, device_index=0):
    import torch
    from metacache_b200 import _lib
    from metacache_b200._lib import Sketching
    from metacache_b200.database import Database
    from metacache_b200.distributed import (DeviceBackend, DeviceReads, FeatureShardedQuery, feature_sharded_step,
                                            load_feature_shard)
    L = _lib.lib()
    dev = torch.device("cuda", device_index)
    torch.cuda.set_device(dev)
    db = Database(device_index, 1)

    def feed(d):
        for p in parts:
            d.load_part_arrays(0, *p, batch=20000)

    load_feature_shard(db, rank, world, n_targets, feed, comm)
    per = (len(reads) + world - 1) // world
    mine = reads[rank * per:(rank + 1) * per]
    sk = Sketching(g1.k, g1.s, g1.w, g1.stride)
    main = torch.cuda.Stream(dev)
    streams = [torch.cuda.Stream(dev) for _ in range(3)]
    with torch.cuda.stream(main):
        dr = DeviceReads(mine, g1.stride, dev)
        ws = _lib.check_ptr(L.mcb200_workspace_create(db._h, max(len(mine), 1), max(dr.n_seqs, 1), dr.n_bases + 64, MAXC, 0))
        backend = DeviceBackend(db, world, chunk, MAXC, dev)
        fq = FeatureShardedQuery(backend, comm, g1.s, MAXC, chunk_queries=chunk, n_slots=3, streams=streams)
        top = torch.zeros((max(len(mine), 1), MAXC, 4), dtype=torch.int32, device=dev)
        feature_sharded_step(fq, ws, dr.q, sk, dr.max_win, top)
        main.synchronize()
        out = top.cpu().numpy().view(np.uint32)[:len(mine)]
    stats = dict(fq.stats, keys=db.key_count(0), values=db.value_count(0), loc_bytes=backend.loc_bytes)
    backend.close()
    L.mcb200_workspace_destroy(ws)
    db.close()
    return [[tuple(int(x) for x in c) for c in row if c[1] > 0] for row in out], stats


@pytest.mark.parametrize("world,chunk", [(2, 64), (3, 1000), (5, 7)])
def test_feature_shards_match_the_reference_per_part_merge(world, chunk):
    g1, g2 = G1(), G2()
    n_targets = len(g1.targets)
    nreads = len(g1.reads) if chunk != 7 else 90
    res = _run_threads(world, lambda comm, r: _rank(comm, r, world, g1.reads[:nreads], chunk, g1, g2.parts, n_targets))
    got = [x for r in res for x in r[0]]
    want = _expected(nreads)
    assert len(got) == nreads
    bad = [i for i in range(nreads) if got[i] != want[i]]
    assert not bad, (bad[:5], got[bad[0]], want[bad[0]])
    # every key of the two parts lives on exactly one shard; locations are all kept
    assert sum(r[1]["values"] for r in res) == sum(len(p[2]) for p in g2.parts)
    assert sum(r[1]["features_sent"] for r in res) > 1000


def test_single_shard_equals_the_plain_query():
    """n_shards = 1: route / probe / gather / reduce with one owner == mcb200_query_device"""
    g1 = G1()
    from metacache_b200.database import Database, query_reads
    res = _run_threads(1, lambda comm, r: _rank(comm, r, 1, g1.reads, 128, g1, [(g1.keys, g1.sizes, g1.values)],
                                                len(g1.targets)))
    db = Database(0, 1)
    db.load_part_arrays(0, g1.keys, g1.sizes, g1.values)
    ref = query_reads(db, g1.reads, copy_all_hits=False)
    db.close()
    for i, (_, top) in enumerate(ref):
        assert res[0][0][i] == top, i


def test_routing_agrees_with_the_restated_owner_function():
    import torch
    from metacache_b200 import _lib
    from metacache_b200.database import Database
    from tests.shard_numpy_backend import shard_of_np
    L = _lib.lib()
    g1 = G1()
    dev = torch.device("cuda", 0)
    db = Database(0, 1)
    db.load_part_arrays(0, g1.keys, g1.sizes, g1.values)
    ws = _lib.check_ptr(L.mcb200_workspace_create(db._h, 16, 16, 64, MAXC, 0))
    rng = np.random.default_rng(3)
    nq, S, N = 9, 16, 5
    wins = rng.integers(0, 4, nq)
    qwo = np.concatenate([[0], np.cumsum(wins)]).astype(np.int32)
    feats = rng.integers(0, 2 ** 32 - 1, (int(qwo[-1]), S), dtype=np.uint64).astype(np.uint32)
    feats[rng.random(feats.shape) < 0.2] = 0xFFFFFFFF
    d_f = torch.from_numpy(feats.view(np.int32)).to(dev)
    d_q = torch.from_numpy(qwo).to(dev)
    pos = torch.zeros(N * (nq + 1) + 1, dtype=torch.int32, device=dev)
    send = torch.zeros(feats.size + 1, dtype=torch.int32, device=dev)
    _lib.check(L.mcb200_shard_route_device(ws, d_f.data_ptr(), d_q.data_ptr(), nq, S, N, pos.data_ptr(), send.data_ptr(), None))
    torch.cuda.synchronize()
    pos, send = pos.cpu().numpy(), send.cpu().numpy().view(np.uint32)
    for q in range(nq):
        f = feats[qwo[q]:qwo[q + 1]].reshape(-1)
        f = f[f != 0xFFFFFFFF]
        o = shard_of_np(f, N)
        for s in range(N):
            b, e = pos[s * (nq + 1) + q], pos[s * (nq + 1) + q + 1]
            assert sorted(send[b:e].tolist()) == sorted(f[o == s].tolist()), (q, s)
    assert pos[-1] == (feats != 0xFFFFFFFF).sum()
    L.mcb200_workspace_destroy(ws)
    db.close()


def test_all_parts_merged_into_one_table_equal_the_reference_per_part_merge():
    """"shard 0 of 1": every feature of both parts of golden g2 in ONE table (buckets concatenated in part
    order, part-major target numbering inside, original ids outside), queried through the ordinary batch API"""
    from metacache_b200.database import Database, query_reads
    from metacache_b200.distributed import ThreadComm, load_feature_shard
    g1, g2 = G1(), G2()
    db = Database(0, 1)

    def feed(d):
        for p in g2.parts:
            d.load_part_arrays(0, *p, batch=30000)

    load_feature_shard(db, 0, 1, len(g1.targets), feed, ThreadComm(ThreadComm.Shared(1), 0))
    assert db.value_count(0) == sum(len(p[2]) for p in g2.parts)
    assert db.key_count(0) == len(np.union1d(g2.parts[0][0], g2.parts[1][0]))
    res = query_reads(db, g1.reads, copy_all_hits=False, batch_queries=500)
    want = _expected(len(g1.reads))
    bad = [i for i, (_, top) in enumerate(res) if top != want[i]]
    assert not bad, (bad[:5], res[bad[0]][1], want[bad[0]])
    db.close()


def test_cpp_shim_merges_the_parts_of_a_database_into_one_table(tmp_path):
    """MCB200_MERGE_PARTS=1: host/shim_query (database_query.hpp:87-124 on the shim) loads the two `.cache`
    files of golden g2 into ONE merged table on one GPU; the target map is sized from the locations
    (MCB200_TARGETS_AUTO); results = the reference's per-part outputs merged in part order"""
    import os
    import subprocess
    from metacache_b200 import dbformat
    from oracle import refio
    g1, g2 = G1(), G2()
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(here, "..", "metacache_b200", "host", "shim_query")
    assert os.path.exists(exe), "run build() first"
    base = str(tmp_path / "g2")
    for p in (0, 1):
        dbformat.write_cache(f"{base}.cache{p}", dbformat.CachePart(*g2.parts[p]))
    rt = str(tmp_path / "reads.txt")
    norm = lambda x: x if len(x) else b"-"
    refio.write_reads_txt(rt, [(norm(a), norm(b)) if len(b) else norm(a) for a, b in g1.reads])
    env = dict(os.environ, MCB200_MERGE_PARTS="1")
    out = subprocess.run([exe, base, rt, "2"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-1500:]
    rows = [ln for ln in out.stdout.splitlines() if not ln.startswith("#")]
    assert "# parts=2 devices=1" in out.stdout and len(rows) == len(g1.reads)
    want = _expected(len(g1.reads))
    for i, ln in enumerate(rows):
        got = [tuple(int(x) for x in c.split(":")) for c in ln.split("\t")[1].split(",") if c]
        assert got == want[i], i
