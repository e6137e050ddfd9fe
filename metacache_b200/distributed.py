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

import numpy as np

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
    """Target-sharded query step (the reference's partitioning): every rank contributes `nq` queries with
    at most `nwin` windows (a capacity: reads of any length fit as long as the rank's window count stays
    below it - the window bound n_bases / winstride + 2 * n_seqs always does; the receivers only read the
    windows the gathered per-read window offsets name)."""

    def __init__(self, db, ws, nq: int, nwin: int, sketchlen: int, max_candidates: int, device, stream,
                 group=None):
        import torch
        import torch.distributed as dist
        self.dist, self.torch = dist, torch
        self.db, self.ws, self.nq, self.nwin, self.S, self.k = db, ws, nq, nwin, sketchlen, max_candidates
        self.device, self.stream, self.group = device, stream, group
        self.comm_stream = torch.cuda.Stream(device)
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
        f_ptr, w_ptr = L.mcb200_workspace_sketches(self.ws), L.mcb200_workspace_query_windows(self.ws)
        mine_f = _as_tensor(f_ptr, self.nwin * self.S, self.device)
        mine_w = _as_tensor(w_ptr, self.nq + 1, self.device)
        sketched = torch.cuda.Event()
        sketched.record(self.stream)
        # the gathers run on their own stream while this rank probes ITS reads straight from the workspace
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(sketched)
            dist.all_gather_into_tensor(self.feats_all, mine_f, group=self.group)
            dist.all_gather_into_tensor(self.qwo_all, mine_w, group=self.group)
            dist.all_gather_into_tensor(self.maxwin_all, max_win, group=self.group)
            gathered = torch.cuda.Event()
            gathered.record(self.comm_stream)
        _lib.check(L.mcb200_query_sketches_device(
            self.ws, 0, f_ptr, w_ptr, max_win.data_ptr(), self.nq, self.S, self.send[self.rank].data_ptr(), self.sp))
        self.stream.wait_event(gathered)
        for j in range(self.world):
            if j == self.rank:
                continue
            _lib.check(L.mcb200_query_sketches_device(
                self.ws, 0, self.feats_all[j].data_ptr(), self.qwo_all[j].data_ptr(),
                self.maxwin_all[j].data_ptr(), self.nq, self.S, self.send[j].data_ptr(), self.sp))
        with torch.cuda.stream(self.stream):
            dist.all_to_all_single(self.recv, self.send, group=self.group)
        _lib.check(L.mcb200_merge_candidates_device(self.ws, self.recv.data_ptr(), self.world, self.nq,
                                                    self.top.data_ptr(), self.sp))
        return self.top


# ======================================================================================
# Feature-space sharding (include/mcb200.h "feature-space sharding", csrc/kernels_shard.cu)
# ======================================================================================
#
# Rank r holds the features f with shard_of(f) = r, each with the locations of ALL database parts.
# One step over this rank's reads, in chunks that are software-pipelined over a few CUDA streams so
# that the exchanges of one chunk overlap the kernels of the next:
#
#   A  route       features of the chunk grouped by owner             (mcb200_shard_route_device)
#      all-gather of the per-owner counts                              -> host (split sizes)
#   B  all-to-all  features to their owners
#      probe       slot lookup + scan of the bucket sizes              (mcb200_shard_probe_device)
#      all-gather of the per-origin location counts                    -> host (split sizes)
#   C  gather      bucket contents, one run per (origin, read)         (mcb200_shard_gather_device)
#      all-to-all  locations + per-feature offsets back to the origins
#      reduce      aggregate / window sums / top candidates            (mcb200_shard_reduce_device)
#
# The host logic (split sizes, segment bookkeeping, pipelining) is independent of where the four
# compute steps run: `backend` supplies them (DeviceBackend = libmcb200 on CUDA tensors; the CPU
# tests inject a numpy backend and run this class over gloo), `comm` supplies the two collectives
# (TorchComm = torch.distributed; ThreadComm = N ranks as threads of one process, for one-GPU tests).

class TorchComm:
    """all-gather of small count vectors (to the host) and variable all-to-all over torch.distributed"""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def all_gather_counts(self, counts):
        """counts: int64 tensor [k] on the compute device -> pending object; .result() = numpy [world, k]"""
        flat = self.torch.empty(self.world * counts.numel(), dtype=counts.dtype, device=counts.device)
        self.dist.all_gather_into_tensor(flat, counts.contiguous().view(-1), group=self.group)
        out = flat.view(self.world, counts.numel())
        host = self.torch.empty(out.shape, dtype=out.dtype, pin_memory=out.is_cuda)
        host.copy_(out, non_blocking=True)
        ev = None
        if out.is_cuda:
            ev = self.torch.cuda.Event()
            ev.record()
        return _Pending(host, ev, out)

    def all_to_all(self, out, inp, out_splits, in_splits, local="copy"):
        """variable all-to-all; the segment a rank sends to itself never enters NCCL: it is copied
        directly (local="copy") or left to the caller, who reads it in place (local="skip")"""
        osp, isp = [int(x) for x in out_splits], [int(x) for x in in_splits]
        r = self.rank
        if self.world == 1:
            if local == "copy":
                out[:osp[0]].copy_(inp[:isp[0]])
            return
        o0, i0 = sum(osp[:r]), sum(isp[:r])
        if local == "copy" and osp[r]:
            out[o0:o0 + osp[r]].copy_(inp[i0:i0 + isp[r]], non_blocking=True)
        # the remote segments: one batch of point-to-point operations (a single NCCL group)
        ops, oo, ii = [], 0, 0
        for j in range(self.world):
            if j != r:
                if isp[j]:
                    ops.append(self.dist.P2POp(self.dist.isend, inp[ii:ii + isp[j]], j, group=self.group))
                if osp[j]:
                    ops.append(self.dist.P2POp(self.dist.irecv, out[oo:oo + osp[j]], j, group=self.group))
            oo += osp[j]
            ii += isp[j]
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()


class _Pending:
    def __init__(self, host, ev, keep):
        self.host, self.ev, self.keep = host, ev, keep

    def result(self):
        if self.ev is not None:
            self.ev.synchronize()
        return self.host.numpy()


class ThreadComm:
    """N ranks as N threads of ONE process (one GPU or the CPU): the same collectives through shared
    memory.  Lets a single-GPU box run the complete sharded data path with several shards."""

    class Shared:
        def __init__(self, world):
            import threading
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, shared, rank):
        import torch
        self.torch, self.sh, self.rank, self.world = torch, shared, rank, shared.world

    def _exchange(self, obj):
        self.sh.slots[self.rank] = obj
        self.sh.barrier.wait()
        got = list(self.sh.slots)
        self.sh.barrier.wait()
        return got

    def all_gather_counts(self, counts):
        if counts.is_cuda:
            self.torch.cuda.current_stream().synchronize()
        rows = self._exchange(counts.cpu())
        return _Pending(self.torch.stack(rows), None, None)

    def all_to_all(self, out, inp, out_splits, in_splits, local="copy"):
        if inp.is_cuda:
            self.torch.cuda.current_stream().synchronize()
        parts = list(self.torch.split(inp, [int(x) for x in in_splits]))
        got = self._exchange(parts)
        o = 0
        for src in range(self.world):
            n = int(out_splits[src])
            out[o:o + n].copy_(got[src][self.rank])
            o += n
        if out.is_cuda:
            self.torch.cuda.current_stream().synchronize()
        self.sh.barrier.wait()                      # nobody reuses `inp` before every reader is done


class DeviceBackend:
    """the four compute steps on CUDA through libmcb200 (one small workspace per pipeline slot)"""

    def __init__(self, db, n_shards, max_chunk_queries, max_candidates, device):
        import torch
        self.torch, self.db, self.N, self.k, self.device = torch, db, n_shards, max_candidates, device
        self.L = _lib.lib()
        self.loc_bytes = self.L.mcb200_db_location_bytes(db._h, 0)
        if self.loc_bytes not in (4, 8):
            raise _lib.Mcb200Error(None, "database part 0 is not loaded")
        self.loc_dtype = torch.int32 if self.loc_bytes == 4 else torch.int64
        self.max_chunk = max_chunk_queries
        self._ws = {}
        self._cap = {}

    def ws(self, slot):
        if slot not in self._ws:
            self._ws[slot] = _lib.check_ptr(self.L.mcb200_workspace_create(
                self.db._h, self.max_chunk, self.max_chunk, 64, self.k, 0))
        return self._ws[slot]

    def close(self):
        for w in self._ws.values():
            self.L.mcb200_workspace_destroy(w)
        self._ws = {}

    @staticmethod
    def _sp():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def empty(self, n, dtype):
        return self.torch.empty(max(int(n), 1), dtype=dtype, device=self.device)

    def route(self, slot, feats_ptr, qwo_ptr, nq, sketchlen, pos, send):
        _lib.check(self.L.mcb200_shard_route_device(self.ws(slot), feats_ptr, qwo_ptr, nq, sketchlen, self.N,
                                                    pos.data_ptr(), send.data_ptr(), self._sp()))

    def probe(self, slot, feats, n, off, data):
        _lib.check(self.L.mcb200_shard_probe_device(self.ws(slot), 0, feats.data_ptr(), n, off.data_ptr(),
                                                    data.data_ptr(), self._sp()))

    def gather(self, slot, off, data, n, locs):
        _lib.check(self.L.mcb200_shard_gather_device(self.ws(slot), 0, off.data_ptr(), data.data_ptr(), n,
                                                     locs.data_ptr(), self._sp()))

    def reduce(self, slot, pos, runs, max_win, nq, top, mean_locations=0.0):
        """runs: list of (locs tensor, off tensor, n_features, n_locations) per owner"""
        # first-pass table of the fused reduction: large databases return many unrelated single hits
        # (distinct locations); for more than two candidates per read (no filter in the kernel) size it by
        # the mean list length instead of overflowing to the second pass
        cap = 256 if mean_locations < 160 else (512 if mean_locations < 320 else 1024)
        if self.k <= 2:
            cap = 256         # the single-hit filter of the kernel keeps unrelated locations out of the table
        if self._cap.get(slot) != cap:
            _lib.check(self.L.mcb200_workspace_set_warp_capacity(self.ws(slot), cap))
            self._cap[slot] = cap
        arr = (_lib.ShardRun * self.N)()
        for o, (locs, off, nf, nl) in enumerate(runs):
            arr[o] = _lib.ShardRun(locs.data_ptr(), off.data_ptr(), int(nf), int(nl))
        _lib.check(self.L.mcb200_shard_reduce_device(self.ws(slot), 0, self.N, pos.data_ptr(), arr,
                                                     max_win.data_ptr(), nq, top.data_ptr(), self._sp()))

    def check(self, slot):
        """-> True if the chunk must be re-issued (a read outgrew the scratch pool, now grown)"""
        rc = self.L.mcb200_workspace_check(self.ws(slot))
        if rc == _lib.EAGAIN:
            return True
        _lib.check(rc)
        return False


class FeatureShardedQuery:
    """One rank of the feature-sharded query.  `sketches` supplies the chunk inputs:
    sketches(q0, q1) -> (feats_ref, qwo_ref, max_win tensor [q1-q0]) where feats_ref/qwo_ref are
    whatever backend.route understands (device pointers for DeviceBackend)."""

    def __init__(self, backend, comm, sketchlen, max_candidates, chunk_queries, n_slots=3, streams=None):
        self.b, self.comm, self.S, self.k = backend, comm, sketchlen, max_candidates
        self.N, self.rank = comm.world, comm.rank
        self.chunk, self.n_slots = chunk_queries, n_slots
        self.streams = streams                     # list of n_slots torch streams (None on the CPU)
        self._bufs = [dict() for _ in range(n_slots)]
        self.stats = {"features_sent": 0, "locations_received": 0, "chunks": 0}
        self.events = None                         # set by enable_timing()

    # ---- helpers ----------------------------------------------------------------------
    def _buf(self, slot, name, n, dtype):
        d = self._bufs[slot]
        t = d.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = self.b.empty(int(n * 1.25) + 16, dtype)
            d[name] = t
        return t

    def _on(self, slot):
        import contextlib
        if self.streams is None:
            return contextlib.nullcontext()
        return self.b.torch.cuda.stream(self.streams[slot])

    def enable_timing(self, on=True):
        self.events = [] if on else None

    def _timed(self, slot, name):
        """context: CUDA events on the slot's stream around the operations enqueued inside"""
        import contextlib
        if self.events is None or self.streams is None:
            return contextlib.nullcontext()
        fq, torch = self, self.b.torch

        class _T:
            def __enter__(self_):
                self_.e0 = torch.cuda.Event(enable_timing=True)
                self_.e0.record(fq.streams[slot])

            def __exit__(self_, *exc):
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record(fq.streams[slot])
                fq.events.append((name, self_.e0, e1))
        return _T()

    def phase_ms(self):
        """sum of the recorded intervals per operation (call after a synchronisation; clears them)"""
        acc = {}
        for name, e0, e1 in self.events or []:
            acc[name] = acc.get(name, 0.0) + e0.elapsed_time(e1)
        if self.events is not None:
            self.events = []
        return acc

    # ---- the three phases of one chunk --------------------------------------------------
    def _phase_a(self, st):
        torch = self.b.torch
        N, nq, slot = self.N, st["nq"], st["slot"]
        with self._on(slot):
            pos = self._buf(slot, "pos", N * (nq + 1) + 1, torch.int32)
            send = self._buf(slot, "send", st["feat_cap"], torch.int32)
            with self._timed(slot, "route"):
                self.b.route(slot, st["feats"], st["qwo"], nq, self.S, pos, send)
            seg_idx = torch.arange(N + 1, dtype=torch.int64, device=pos.device) * (nq + 1)
            seg = pos[seg_idx].to(torch.int64)                  # N + 1 segment starts
            st["pos"], st["send"] = pos, send
            st["pend_a"] = self.comm.all_gather_counts(seg[1:] - seg[:-1])

    def _phase_b(self, st):
        torch = self.b.torch
        N, slot = self.N, st["slot"]
        pend = st.pop("pend_a")
        M = pend.result()                                       # M[g][o]: features g sends to o
        st["send_counts"] = M[self.rank].copy()
        st["recv_counts"] = M[:, self.rank].copy()
        nrecv = int(st["recv_counts"].sum())
        st["nrecv"] = nrecv
        with self._on(slot):
            recv = self._buf(slot, "recv", nrecv, torch.int32)
            with self._timed(slot, "exchange_features"):
                self.comm.all_to_all(recv[:nrecv], st["send"][:int(st["send_counts"].sum())],
                                     st["recv_counts"], st["send_counts"])
            off = self._buf(slot, "off", nrecv + 1, torch.int32)
            data = self._buf(slot, "data", nrecv, torch.int64)
            with self._timed(slot, "probe"):
                self.b.probe(slot, recv, nrecv, off, data)
            # location offsets at the origin boundaries, computed where the counts already are (no host copy)
            if pend.keep is not None:
                rc = pend.keep[:, self.rank]
            else:
                rc = torch.as_tensor(st["recv_counts"], dtype=torch.int64).to(off.device)
            bounds = torch.cat([torch.zeros(1, dtype=torch.int64, device=off.device), torch.cumsum(rc, 0)])
            segoff = off[bounds].to(torch.int64)
            st["off"], st["data"] = off, data
            st["pend_b"] = self.comm.all_gather_counts(segoff[1:] - segoff[:-1])
        self.stats["features_sent"] += int(st["send_counts"].sum())

    def _phase_c(self, st, top):
        torch = self.b.torch
        N, slot, nq = self.N, st["slot"], st["nq"]
        Lc = st.pop("pend_b").result()                          # Lc[o][g]: locations owner o returns to origin g
        loc_send = Lc[self.rank].copy()
        loc_recv = Lc[:, self.rank].copy()
        nsend, nrecv_l = int(loc_send.sum()), int(loc_recv.sum())
        if max(nsend, nrecv_l) >= 2 ** 32 - 2:
            raise _lib.Mcb200Error(None, "more than 2^32 locations in one chunk: use smaller chunks")
        with self._on(slot):
            locs = self._buf(slot, "locs", nsend, self.b.loc_dtype)
            with self._timed(slot, "gather"):
                self.b.gather(slot, st["off"], st["data"], st["nrecv"], locs)
            rlocs = self._buf(slot, "rlocs", nrecv_l, self.b.loc_dtype)
            nfs = int(st["send_counts"].sum())
            roff = self._buf(slot, "roff", nfs, torch.int32)
            with self._timed(slot, "exchange_locations"):
                # what this rank returns to itself is read in place by the reduction (no copy, no NCCL)
                self.comm.all_to_all(rlocs[:nrecv_l], locs[:nsend], loc_recv, loc_send, local="skip")
                self.comm.all_to_all(roff[:nfs], st["off"][:st["nrecv"]], st["send_counts"], st["recv_counts"],
                                     local="skip")
            runs, fo, lo = [], 0, 0
            me = self.rank
            own_l0, own_f0 = int(loc_send[:me].sum()), int(st["recv_counts"][:me].sum())
            for o in range(N):
                nf, nl = int(st["send_counts"][o]), int(loc_recv[o])
                if o == me:
                    runs.append((locs[own_l0:own_l0 + max(nl, 1)], st["off"][own_f0:own_f0 + max(nf, 1)], nf, nl))
                else:
                    runs.append((rlocs[lo:lo + max(nl, 1)], roff[fo:fo + max(nf, 1)], nf, nl))
                fo += nf
                lo += nl
            with self._timed(slot, "reduce"):
                self.b.reduce(slot, st["pos"], runs, st["max_win"], nq, top[st["q0"]:st["q0"] + nq],
                              mean_locations=nrecv_l / max(nq, 1))
        self.stats["locations_received"] += nrecv_l
        self.stats["chunks"] += 1

    # ---- one step ----------------------------------------------------------------------
    def step(self, nq_total, sketches, top, feat_cap):
        """All reads of this rank: nq_total reads, top = output tensor [nq_total, k, 4] (int32 view
        of mcb200_candidate).  Every rank must call step() the same number of times; ranks may hold
        different numbers of reads (the chunk count is agreed on first)."""
        torch = self.b.torch
        nchunks = (nq_total + self.chunk - 1) // self.chunk
        agreed = self.comm.all_gather_counts(torch.tensor([nchunks], dtype=torch.int64, device=top.device)).result()
        nchunks = int(agreed.max())
        states = {}

        def make(c):
            q0 = min(c * self.chunk, nq_total)
            q1 = min(q0 + self.chunk, nq_total)
            feats, qwo, max_win = sketches(q0, q1)
            nq = q1 - q0
            return {"c": c, "slot": c % self.n_slots, "q0": q0, "nq": nq, "feats": feats, "qwo": qwo,
                    "max_win": max_win, "feat_cap": feat_cap}

        for t in range(nchunks + 2):
            if t < nchunks:
                states[t] = make(t)
                self._phase_a(states[t])
            if 0 <= t - 1 < nchunks:
                self._phase_b(states[t - 1])
            if 0 <= t - 2 < nchunks:
                self._phase_c(states.pop(t - 2), top)
        return top


class DeviceReads:
    """A list of reads (bytes or (mate1, mate2)) as the device arrays of mcb200_dev_queries, with the
    candidate rules of make_candidate_generation_rules (candidate_structs.hpp:134-151)."""

    def __init__(self, reads, winstride, device, insert_size_max=0):
        import torch
        seqs, seq_query, mw = [], [], []
        for qi, r in enumerate(reads):
            a, b = (r, b"") if isinstance(r, (bytes, bytearray)) else r
            for m in ([m for m in (a, b) if len(m)] or [b""]):
                seqs.append(m)
                seq_query.append(qi)
            mw.append(2 + max(len(a) + len(b), insert_size_max) // winstride)
        offs = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
        flat = np.frombuffer(b"".join(seqs) + b"\0" * 64, np.uint8)
        self.n_queries, self.n_seqs, self.n_bases = len(reads), len(seqs), int(offs[-1])
        self.bases = torch.from_numpy(flat.copy()).to(device)
        self.seq_off = torch.from_numpy(offs.astype(np.uint32).view(np.int32).copy()).to(device)
        self.seq_query = torch.from_numpy(np.array(seq_query, np.int32)).to(device)
        self.max_win = torch.from_numpy(np.array(mw, np.int32)).to(device)
        self.q = DevQueries(self.bases.data_ptr(), self.seq_off.data_ptr(), self.seq_query.data_ptr(),
                            self.max_win.data_ptr(), self.n_seqs, self.n_queries, self.n_bases)


def feature_sharded_step(fq, ws, reads_q, sk, max_win, top, n_seqs_bound=None):
    """One step of FeatureShardedQuery on the device for the reads of `reads_q` (DevQueries): sketches
    them all with the rank's full-batch workspace `ws` on the current stream, then runs the chunk
    pipeline.  Re-issues the step if a read outgrew a scratch pool.  -> top"""
    import torch
    L = _lib.lib()
    cur = torch.cuda.current_stream()
    _lib.check(L.mcb200_sketch_device(ws, C.byref(reads_q), C.byref(sk), C.c_void_p(cur.cuda_stream)))
    done = torch.cuda.Event()
    done.record(cur)
    if fq.streams is not None:
        for s in fq.streams:
            s.wait_event(done)
    feats_ptr = L.mcb200_workspace_sketches(ws)
    qwo_ptr = L.mcb200_workspace_query_windows(ws)
    nq = reads_q.n_queries
    feat_cap = (reads_q.n_bases // sk.winstride + 2 * reads_q.n_seqs) * sk.sketchlen + 64

    def sketches(q0, q1):
        return feats_ptr, qwo_ptr + 4 * q0, max_win[q0:q1]

    for attempt in range(6):
        fq.step(nq, sketches, top, feat_cap)
        if fq.streams is not None:
            for s in fq.streams:
                cur.wait_stream(s)
        again = False
        for slot in range(fq.n_slots):
            if slot in fq.b._ws:
                again |= fq.b.check(slot)
        # every rank must take the same decision
        flag = fq.comm.all_gather_counts(torch.tensor([int(again)], dtype=torch.int64, device=top.device)).result()
        if not flag.any():
            return top
    raise _lib.Mcb200Error(_lib.EAGAIN, "scratch pools still too small after 6 attempts")


def load_feature_shard(db, shard, n_shards, n_targets, feed_parts, comm, max_load_factor=0.0):
    """Sharded load of part slot 0 of `db`: feed_parts(db) must push EVERY part of the database, in part
    order, through the usual loaders (Database.load_part_arrays(0, ...), mcb200_db_load_cache_file,
    mcb200_db_build_part_from_targets on slot 0).  The shards agree on one location packing."""
    import torch
    L = _lib.lib()
    dev = torch.device("cuda", db.device) if torch.cuda.is_available() else torch.device("cpu")
    mt, mw = C.c_uint32(0), C.c_uint32(0)
    # a rank that fails (out of memory while collecting, say) must not leave the others waiting in a
    # collective: every rank reports its status with the maxima, and all raise together
    err = None
    try:
        _lib.check(L.mcb200_db_shard_begin(db._h, 0, shard, n_shards, n_targets))
        feed_parts(db)
        _lib.check(L.mcb200_db_shard_maxima(db._h, 0, C.byref(mt), C.byref(mw)))
    except Exception as ex:                              # noqa: BLE001 - re-raised on every rank below
        err = ex
    mx = comm.all_gather_counts(torch.tensor([mt.value, mw.value, int(err is None)], dtype=torch.int64, device=dev)).result()
    if err is not None or not mx[:, 2].all():
        raise RuntimeError(f"sharded load failed on rank(s) {[int(r) for r in np.flatnonzero(mx[:, 2] == 0)]}: {err}")
    try:
        _lib.check(L.mcb200_db_shard_finish(db._h, 0, max_load_factor, int(mx[:, 0].max()), int(mx[:, 1].max())))
    except Exception as ex:                              # noqa: BLE001
        err = ex
    okf = comm.all_gather_counts(torch.tensor([int(err is None)], dtype=torch.int64, device=dev)).result()
    if err is not None or not okf.all():
        raise RuntimeError(f"building the shard table failed on rank(s) {[int(r) for r in np.flatnonzero(okf[:, 0] == 0)]}: {err}")
