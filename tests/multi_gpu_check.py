"""Run under torchrun on N >= 2 GPUs: checks the NCCL-sharded query (ShardedQuery) against a
single-process N-part database on rank 0 (whose merge path is pinned to the reference by
tests/test_gpu_parity.py::test_two_parts_merge_matches_reference_per_part).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
         --master-port 29533 tests/multi_gpu_check.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metacache_b200 import _lib, synth  # noqa: E402
from metacache_b200._lib import DevQueries, Sketching  # noqa: E402
from metacache_b200.database import Database  # noqa: E402
from metacache_b200.distributed import ShardedQuery  # noqa: E402

NT, TL, NQ, RL, MAXC = 300, 20000, 20000, 150, 2


def build(db, part_slot, part, device):
    bases, off = synth.make_targets(NT, TL, 10, synth.SEED_DB + part, device=device)
    sk = Sketching(16, 16, 127, 112)
    _lib.check(_lib.lib().mcb200_db_build_part_from_targets(db._h, part_slot, bases.data_ptr(), off.data_ptr(), NT,
                                                            part * NT, C.byref(sk), 254, 0.0, None))
    return bases


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=device)
    L = _lib.lib()
    sk = Sketching(16, 16, 127, 112)
    db = Database(device.index, 1)
    bases = build(db, 0, rank, device)
    reads = synth.make_reads_150(NQ, bases, NT, TL, RL, seed=synth.SEED_R150 + 1000 * rank, device=device)
    flat = reads.reshape(-1)
    seq_off = (torch.arange(NQ + 1, dtype=torch.int64, device=device) * RL).to(torch.int32)
    seq_qry = torch.arange(NQ, dtype=torch.int32, device=device)
    max_win = torch.full((NQ,), 2 + RL // 112, dtype=torch.int32, device=device)
    stream = torch.cuda.Stream(device)
    ws = _lib.check_ptr(L.mcb200_workspace_create(db._h, NQ, NQ, NQ * RL + 64, MAXC, 0))
    q = DevQueries(flat.data_ptr(), seq_off.data_ptr(), seq_qry.data_ptr(), max_win.data_ptr(), NQ, NQ, NQ * RL)
    sq = ShardedQuery(db, ws, NQ, 2 * NQ, 16, MAXC, device, stream)
    top = sq.step(q, sk, max_win)
    torch.cuda.synchronize(device)
    # gather every rank's reads and results on rank 0
    all_reads = [torch.empty_like(flat) for _ in range(world)]
    all_tops = [torch.empty_like(top) for _ in range(world)]
    dist.all_gather(all_reads, flat)
    dist.all_gather(all_tops, top.contiguous())
    ok = True
    if rank == 0:
        ref = Database(device.index, world)
        for p in range(world):
            build(ref, p, p, device)
        ws2 = _lib.check_ptr(L.mcb200_workspace_create(ref._h, NQ, NQ, NQ * RL + 64, MAXC, 0))
        out = torch.empty((NQ, MAXC, 4), dtype=torch.int32, device=device)
        nonempty = 0
        for j in range(world):
            fj = all_reads[j].contiguous()
            qj = DevQueries(fj.data_ptr(), seq_off.data_ptr(), seq_qry.data_ptr(), max_win.data_ptr(), NQ, NQ, NQ * RL)
            _lib.check(L.mcb200_query_device(ws2, C.byref(qj), C.byref(sk), out.data_ptr(), None))
            torch.cuda.synchronize(device)
            same = bool(torch.equal(out, all_tops[j]))
            nonempty += int((out[:, 0, 1] > 0).sum())
            print(f"slice {j}: sharded == single-process {world}-part: {same}")
            ok &= same
        ok &= nonempty > NQ // 2
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", f"({world} ranks, {nonempty} reads with hits)")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
