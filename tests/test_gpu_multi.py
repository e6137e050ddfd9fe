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
