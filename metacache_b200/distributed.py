"""Database sharded by target over the GPUs of one box, one process per GPU.

Rank r holds database part r (the reference's own partitioning: whole targets per part,
`building.cpp:352-380`; one part per GPU, `gpu_hashmap.cu:1320-1362`).  One step:

  1. every rank sketches ITS slice of the reads            (mcb200_sketch_device)
  2. NCCL all-gather of the sketches (64 B / window)        -> every rank has every sketch
  3. every rank probes ALL reads against its part            (mcb200_query_sketches_device)
  4. NCCL all-to-all of the partial top hits by read slice   (16*k B / read / part)
  5. stable part-ordered merge on the device                 (mcb200_merge_candidates_device)

The reference instead chains the GPUs with cudaMemcpyPeerAsync and lets GPU 0 sketch alone
(`query_batch.cu:464-527, 646-652`).  torch.distributed is the plumbing; all compute is in
libmcb200.so.  Slices must be equally sized (pad the last one).
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import DevQueries, Sketching


def _as_tensor(ptr, n, device):
    """int32 torch view of `n` u32 at device pointer `ptr` (library-owned memory)"""
    import torch

    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device=device)


class ShardedQuery:
    """Fixed-shape sharded query step: every rank contributes `nq` queries with `nwin` windows."""

    def __init__(self, db, ws, nq: int, nwin: int, sketchlen: int, max_candidates: int, device, stream,
                 group=None):
        import torch
        import torch.distributed as dist
        self.dist, self.torch = dist, torch
        self.db, self.ws, self.nq, self.nwin, self.S, self.k = db, ws, nq, nwin, sketchlen, max_candidates
        self.device, self.stream, self.group = device, stream, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        w = self.world
        self.feats_all = torch.empty((w, nwin * sketchlen), dtype=torch.int32, device=device)
        self.qwo_all = torch.empty((w, nq + 1), dtype=torch.int32, device=device)
        self.maxwin_all = torch.empty((w, nq), dtype=torch.int32, device=device)
        self.send = torch.empty((w, nq, max_candidates, 4), dtype=torch.int32, device=device)
        self.recv = torch.empty((w, nq, max_candidates, 4), dtype=torch.int32, device=device)
        self.top = torch.empty((nq, max_candidates, 4), dtype=torch.int32, device=device)
        self.sp = C.c_void_p(stream.cuda_stream)

    def step(self, q: DevQueries, sk: Sketching, max_win):
        """q: this rank's reads (device); max_win: int32 tensor [nq].  Returns self.top:
        final candidates of THIS rank's slice."""
        L, dist, torch = _lib.lib(), self.dist, self.torch
        _lib.check(L.mcb200_sketch_device(self.ws, C.byref(q), C.byref(sk), self.sp))
        mine_f = _as_tensor(L.mcb200_workspace_sketches(self.ws), self.nwin * self.S, self.device)
        mine_w = _as_tensor(L.mcb200_workspace_query_windows(self.ws), self.nq + 1, self.device)
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(self.feats_all, mine_f, group=self.group)
            dist.all_gather_into_tensor(self.qwo_all, mine_w, group=self.group)
            dist.all_gather_into_tensor(self.maxwin_all, max_win, group=self.group)
        for j in range(self.world):
            _lib.check(L.mcb200_query_sketches_device(
                self.ws, 0, self.feats_all[j].data_ptr(), self.qwo_all[j].data_ptr(),
                self.maxwin_all[j].data_ptr(), self.nq, self.S, self.send[j].data_ptr(), self.sp))
        with torch.cuda.stream(self.stream):
            dist.all_to_all_single(self.recv, self.send, group=self.group)
        _lib.check(L.mcb200_merge_candidates_device(self.ws, self.recv.data_ptr(), self.world, self.nq,
                                                    self.top.data_ptr(), self.sp))
        return self.top
