"""The feature-sharded query over NCCL, one process per GPU (needs >= 2 devices; skipped otherwise).
Expected values are the REFERENCE's per-part outputs (golden g2: 2-part database built and queried by
the reference, `part=0` / `part=1` harness runs) merged in part order (docs/partitioning.md:116-142) -
not this repository's own single-process result."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MAXC = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nreads, chunk, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from metacache_b200.distributed import TorchComm
    from tests.golden_util import G1, G2
    from tests.test_gpu_shard import _rank
    g1, g2 = G1(), G2()
    got, stats = _rank(TorchComm(), rank, world, g1.reads[:nreads], chunk, g1, g2.parts, len(g1.targets), device_index=rank)
    torch.save((got, stats), os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _device_count():
    from metacache_b200 import _lib
    return _lib.lib().mcb200_device_count()


@pytest.mark.parametrize("world,chunk", [(2, 100), (4, 1000), (8, 40)])
def test_nccl_ranks_match_the_reference_per_part_merge(world, chunk, tmp_path):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch
    import torch.multiprocessing as mp
    from tests.golden_util import G1
    from tests.test_gpu_shard import _expected
    nreads = len(G1().reads)
    mp.spawn(_worker, args=(world, _free_port(), nreads, chunk, str(tmp_path)), nprocs=world, join=True)
    got = []
    for r in range(world):
        got += torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))[0]
    want = _expected(nreads)
    assert len(got) == nreads
    bad = [i for i in range(nreads) if got[i] != want[i]]
    assert not bad, (bad[:5], got[bad[0]], want[bad[0]])


def test_one_process_store_over_two_devices_matches_the_reference_per_part_merge():
    """mcb200_db_open_multi: part 0 on cuda:0, part 1 on cuda:1, ONE process (the reference's own
    multi-GPU mode, gpu_hashmap.cu:1255-1292): batch API == reference per-part outputs merged in part order"""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from metacache_b200.database import Database, query_reads
    from oracle import mc_oracle as O
    from tests.golden_util import G1, G2
    g1, g2 = G1(), G2()
    for devices in ([0, 1], [1, 0]):
        db = Database(devices=devices)
        for p in (0, 1):
            db.load_part_arrays(p, *g2.parts[p])
        assert [db.key_count(p) for p in (0, 1)] == [len(g2.parts[p][0]) for p in (0, 1)]
        res = query_reads(db, g1.reads, copy_all_hits=False, batch_queries=300)
        e0, e1 = g2.expected(0), g2.expected(1)
        for i, (_, top) in enumerate(res):
            assert top == O.merge_tops([e0.top[i], e1.top[i]], 2), (devices, i)
        db.close()


def test_cpp_shim_spreads_a_two_part_database_over_two_devices(tmp_path):
    """host/shim_query (database_query.hpp:87-124 written against the shim) on a 2-part database:
    prepare_query_tables puts one part on each GPU, in one process"""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    from metacache_b200 import dbformat
    from oracle import mc_oracle as O, refio
    from tests.golden_util import G1, G2
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
    out = subprocess.run([exe, base, rt, "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-1500:]
    e0, e1 = g2.expected(0), g2.expected(1)
    rows = [ln for ln in out.stdout.splitlines() if not ln.startswith("#")]
    assert len(rows) == len(g1.reads)
    assert "# parts=2 devices=2" in out.stdout
    for i, ln in enumerate(rows):
        got = [tuple(int(x) for x in c.split(":")) for c in ln.split("\t")[1].split(",") if c]
        assert got == O.merge_tops([e0.top[i], e1.top[i]], 2), i


def _target_worker(rank, world, port, out_dir):
    import ctypes as C
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from metacache_b200 import _lib
    from metacache_b200._lib import Sketching
    from metacache_b200.database import Database
    from metacache_b200.distributed import DeviceReads, ShardedQuery
    from tests.golden_util import G1, G2
    L = _lib.lib()
    g1, g2 = G1(), G2()
    db = Database(rank, 1)
    db.load_part_arrays(0, *g2.parts[rank])                     # rank r holds part r
    per = (len(g1.reads) + world - 1) // world
    mine = g1.reads[rank * per:(rank + 1) * per]
    mine = mine + [(b"", b"")] * (per - len(mine))              # equal slices: pad with empty reads
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        dr = DeviceReads(mine, g1.stride, dev)
        cap = torch.tensor([dr.n_bases // g1.stride + 2 * dr.n_seqs], dtype=torch.int64, device=dev)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX)              # ragged reads: one window capacity for all ranks
        nwin = int(cap.item())
        ws = _lib.check_ptr(L.mcb200_workspace_create(db._h, per, dr.n_seqs, dr.n_bases + 64, MAXC, 0))
        sq = ShardedQuery(db, ws, per, nwin, g1.s, MAXC, dev, stream)
        for attempt in range(6):                                # re-issued while a scratch pool had to grow (the
            top = sq.step(dr.q, Sketching(g1.k, g1.s, g1.w, g1.stride), dr.max_win)    # 19 kbp tandem-repeat read)
            stream.synchronize()
            rc = L.mcb200_workspace_check(ws)
            assert rc in (0, _lib.EAGAIN), rc
            again = torch.tensor([int(rc != 0)], dtype=torch.int64, device=dev)
            dist.all_reduce(again, op=dist.ReduceOp.MAX)        # every rank takes the same decision
            if attempt > 0 and not int(again.item()):           # at least twice: buffers are reused across steps
                break
        out = top.cpu().numpy().view(np.uint32)
    got = [[tuple(int(x) for x in c) for c in row if c[1] > 0] for row in out]
    torch.save(got, os.path.join(out_dir, f"t{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_target_sharded_nccl_step_matches_the_reference_per_part_merge(tmp_path):
    """ShardedQuery (all-gather of sketches, every rank probes every read against ITS part, all-to-all of
    partial candidates, part-ordered merge) on 2 GPUs, ragged reads, vs the reference's per-part outputs"""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch
    import torch.multiprocessing as mp
    from tests.golden_util import G1
    from tests.test_gpu_shard import _expected
    n = len(G1().reads)
    mp.spawn(_target_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = []
    for r in range(2):
        got += torch.load(os.path.join(str(tmp_path), f"t{r}.pt"))
    want = _expected(n)
    bad = [i for i in range(n) if got[i] != want[i]]
    assert not bad, (bad[:5], got[bad[0]], want[bad[0]])
