#!/usr/bin/env python
"""Benchmark of the MetaCache query hot path on B200 (BASELINE.json metric: reads/s, 150 bp).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (encode -> sketch -> probe -> per-read sort ->
contiguous-window top hits [-> partial-top exchange + merge at N > 1]) over one batch of
synthetic reads.  Workload at N = 1 is BASELINE config C2: 10 M x 150 bp reads (R150 recipe)
against the 50 k-target / 16-mer synthetic database DB-S (metacache_b200/synth.py).  At N > 1
every rank adds one DB-S sized part and 10 M reads (weak scaling).  --shard-by auto (default): all
parts merged into one table on every GPU when that fits its 180 GB (reads shard, no exchange in the
data path), else the reference's partitioning, one part per GPU, with partial top hits exchanged
over NCCL and merged on device (--shard-by target); --shard-by feature: feature-space sharding.

Prints ONE JSON line (see DESIGN.md "Measurement" for the fields).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

# The e2e path keeps 3 batch slots per worker thread in flight, one CUDA stream each.  The driver maps streams
# onto CUDA_DEVICE_MAX_CONNECTIONS hardware channels (default 8): streams sharing a channel serialise.  Measured
# on C2 with 16 workers: 380 -> 422 M reads/s end to end with 32 channels (profiles/README.md).  The library
# sets the same default when it is loaded before the CUDA context exists; here torch may come first.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 150
SK = dict(kmerlen=16, sketchlen=16, winlen=127, winstride=112)
MAXC = 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C2", "C3"],
                    help="C2: 150 bp reads (the metric's configuration, default); C3: long reads (200-19000 bp)")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (0 = 10 M for C2, 2 M for C3)")
    ap.add_argument("--targets", type=int, default=50_000, help="targets per database part")
    ap.add_argument("--target-len", type=int, default=100_000)
    ap.add_argument("--slot-reads", type=int, default=125_000, help="reads per host batch slot (e2e)")
    ap.add_argument("--e2e-threads", type=int, default=0, help="host worker threads of the e2e run (0 = one per core, max 32)")
    ap.add_argument("--e2e-slots-per-worker", type=int, default=3, help="batch slots each e2e worker rotates through (one being filled, the others in flight)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end measurement (kernel experiments only)")
    ap.add_argument("--prepare-reference", action="store_true",
                    help="internal (child of --impl reference): write part 0 in the reference's format + the read sample, exit")
    ap.add_argument("--replicate", action="store_true",
                    help="N > 1: every GPU holds the SAME single-part database and queries only its own reads "
                         "(the reference's -replicate mode, SURVEY 8e) instead of the target-sharded database")
    ap.add_argument("--replicate-merged", action="store_true", help="same as --shard-by merged")
    ap.add_argument("--shard-by", default="auto", choices=["auto", "merged", "target", "feature"],
                    help="N > 1, how the N-part database is held (DESIGN.md 5).  merged: every GPU holds ALL parts merged "
                         "into one table (buckets of a feature concatenated in part order) and queries only ITS reads - "
                         "reads shard, nothing is exchanged in the data path; needs the whole database in one GPU's "
                         "180 GB.  target: the reference's partitioning, one part per GPU, every GPU probes every read, "
                         "NCCL all-gather of sketches + all-to-all of partial top hits (the capacity mode).  feature: "
                         "every GPU owns a slice of the feature space, features and location lists travel over NCCL.  "
                         "auto (default): merged if the database fits, else target")
    ap.add_argument("--no-compare-target", action="store_true",
                    help="merged mode at N > 1: skip the short target-sharded measurement reported beside it")
    ap.add_argument("--merged-parts", type=int, default=0,
                    help="merged mode: number of database parts merged on every GPU (0 = one per rank); lets one GPU "
                         "measure the kernel on the N-part database")
    ap.add_argument("--shard-streams", type=int, default=3, help="CUDA streams of the feature-sharded chunk pipeline (1 = every operation serial, for profiling)")
    ap.add_argument("--chunk-reads", type=int, default=1_250_000, help="reads per pipeline chunk of the feature-sharded step")
    ap.add_argument("--table-slots", type=int, default=0, help="per-warp aggregation table slots (0 = library default)")
    ap.add_argument("--load-factor", type=float, default=0.0, help="table load factor (0 = library default)")
    ap.add_argument("--cache", default=os.environ.get(
        "MCB200_BENCH_CACHE", "/dev/shm/mcb200_bench" if os.path.isdir("/dev/shm") else "/tmp/mcb200_bench"))
    return ap.parse_args()


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="mcb200_clocks_", suffix=".csv")
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            busy = [x for x in sm if x >= 0.5 * max(sm)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), samples=len(busy), samples_total=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# --------------------------------------------------------------------------------------
# synthetic workload
# --------------------------------------------------------------------------------------
def db_cache_dir(args, part):
    return os.path.join(args.cache, f"dbs_t{args.targets}_l{args.target_len}_p{part}")


def build_part(args, part, device):
    """DB-S part `part` built on the GPU from synthetic targets.  Returns Database."""
    import torch
    from metacache_b200 import _lib, synth
    from metacache_b200.database import Database
    from metacache_b200._lib import Sketching
    t0 = time.time()
    bases, off = synth.make_targets(args.targets, args.target_len, 10, synth.SEED_DB + part, device=device)
    db = Database(device.index, 1)
    sk = Sketching(**SK)
    wins = np.zeros(args.targets, np.uint32)
    _lib.check(_lib.lib().mcb200_db_build_part_from_targets(
        db._h, 0, bases.data_ptr(), off.data_ptr(), args.targets, part * args.targets, C.byref(sk), 254,
        args.load_factor, wins.ctypes.data))
    torch.cuda.synchronize(device)
    info = dict(build_s=round(time.time() - t0, 2), keys=db.key_count(0), locations=db.value_count(0),
                table_gb=round(db.device_bytes(0) / 1e9, 2))
    return db, bases, wins, info


def build_feature_shard(args, rank, world, device, all_features=False):
    """Shard `rank` of the feature-sharded database: every rank generates and sketches EVERY part
    (no exchange at load time; ~1 s per part) and keeps the features it owns, with the locations of
    all parts merged per feature.  Returns the bases of part `rank` for the read generator."""
    import torch
    from metacache_b200 import _lib, synth
    from metacache_b200.database import Database
    from metacache_b200.distributed import ThreadComm, TorchComm, load_feature_shard
    from metacache_b200._lib import Sketching
    t0 = time.time()
    nparts = max(world, args.merged_parts) if all_features else world
    comm = TorchComm() if world > 1 else ThreadComm(ThreadComm.Shared(1), 0)
    # the merged table is built with room to spare (180 GB): short probe sequences
    lf = args.load_factor or (0.3 if all_features else 0.0)
    db = Database(device.index, 1)
    sk = Sketching(**SK)
    keep = {}

    def feed(d):
        for p in range(nparts):
            bases, off = synth.make_targets(args.targets, args.target_len, 10, synth.SEED_DB + p, device=device)
            wins = np.zeros(args.targets, np.uint32)
            _lib.check(_lib.lib().mcb200_db_build_part_from_targets(
                d._h, 0, bases.data_ptr(), off.data_ptr(), args.targets, p * args.targets, C.byref(sk), 254,
                args.load_factor, wins.ctypes.data))
            if p == rank:
                keep["bases"], keep["wins"] = bases, wins
            del bases, off
            torch.cuda.empty_cache()

    # all_features: "shard 0 of 1" = every feature, i.e. the whole N-part database merged on this GPU
    load_feature_shard(db, 0 if all_features else rank, 1 if all_features else world, nparts * args.targets, feed,
                       comm, lf)
    torch.cuda.synchronize(device)
    info = dict(build_s=round(time.time() - t0, 2), keys=db.key_count(0), locations=db.value_count(0),
                table_gb=round(db.device_bytes(0) / 1e9, 2),
                shard=("all %d parts merged on every GPU" if all_features else "features of all %d parts owned by this rank") % nparts)
    return db, keep["bases"], keep["wins"], info


def make_reads(args, bases, rank, device):
    """-> (flat uint8 bases on the device, int64 offsets [n+1] on the device)"""
    import torch
    from metacache_b200 import synth
    if args.workload == "C3":
        return synth.make_long_reads(args.reads, bases, args.targets, args.target_len,
                                     seed=synth.SEED_RLONG + 1000 * rank, device=device)
    # every rank owns a different slice of the global read set (seed offset by rank); reads are
    # sampled from the rank's own part (the other parts see them as mostly-missing queries,
    # plus whatever the shared 16-mer space yields)
    reads = synth.make_reads_150(args.reads, bases, args.targets, args.target_len, READ_LEN,
                                 seed=synth.SEED_R150 + 1000 * rank, device=device)
    return reads.reshape(-1), torch.arange(args.reads + 1, dtype=torch.int64, device=device) * READ_LEN


def export_reference_db(args, db, wins, part):
    """writes <cache>/db.meta + db.cache0 in the reference's on-disk format"""
    from metacache_b200 import dbformat
    d = db_cache_dir(args, part)
    os.makedirs(d, exist_ok=True)
    base = os.path.join(d, "db")
    if os.path.exists(base + ".meta") and os.path.exists(base + ".cache0") and os.path.exists(base + ".ok"):
        return base
    keys, sizes, values = db.export_part(0)
    if part:
        values = values - (np.uint64(part * args.targets) << np.uint64(32))    # local target ids
    dbformat.write_cache(base + ".cache0", dbformat.CachePart(keys, sizes, values))
    dbformat.write_meta(base + ".meta", dbformat.synthetic_meta(wins, **SK))
    open(base + ".ok", "w").write("ok")
    return base


def write_reads_txt(path, flat_np, offs_np):
    """one read per line"""
    n = len(offs_np) - 1
    lens = np.diff(offs_np)
    out = np.full(int(offs_np[-1] - offs_np[0]) + n, ord("\n"), np.uint8)
    dst = (np.arange(len(flat_np), dtype=np.int64) + np.repeat(np.arange(n, dtype=np.int64), lens))
    out[dst] = flat_np
    out.tofile(path)


def reference_reads_path(args, base, sample):
    return base + f".{args.workload}.first{sample}.txt"


def cpu_reference_run(base, flat_np, offs_np, threads, passes, tops=None, reads_txt=None):
    """the reference's own hot path (oracle/_ref/mc_ref_harness links the unmodified reference
    objects) on `threads` host threads; returns dict with per-pass seconds.  tops: file that
    receives the reference's top candidates of every read ([n][MAXC][4] u32) for the parity check"""
    from oracle import refio
    if not os.path.exists(refio.HARNESS):
        return None
    rt = reads_txt or base + f".reads{len(offs_np) - 1}_{int(offs_np[-1] - offs_np[0])}.txt"
    if not os.path.exists(rt):
        write_reads_txt(rt + ".tmp", flat_np, offs_np)
        os.replace(rt + ".tmp", rt)
    kw = dict(threads=threads, repeat=passes, sketches=0, allhits=0, maxcand=MAXC)
    if tops:
        kw["tops"] = tops
    return refio.run_harness(base, rt, "-", **kw)


def parity_against_reference(tops_file, gpu_top):
    """bit-exact comparison of the reference's top candidates (harness dump) with the rows the GPU
    produced for the same reads: {tgt, hits, beg, end} x MAXC, unused entries {~0, 0, 0, 0}"""
    ref = np.fromfile(tops_file, dtype="<u4").reshape(-1, MAXC, 4)
    got = np.ascontiguousarray(gpu_top[:len(ref)]).view(np.uint32).reshape(-1, MAXC, 4)
    bad = np.flatnonzero((ref != got).any(axis=(1, 2)))
    out = {"reads": int(len(ref)), "mismatches": int(len(bad)), "reads_with_hits": int((ref[:, 0, 1] > 0).sum()),
           "against": "reference database::query_host top candidates (oracle/_ref/mc_ref_harness), same database file, same reads"}
    if len(bad):
        out["first_mismatch"] = {"read": int(bad[0]), "reference": ref[bad[0]].tolist(), "gpu": got[bad[0]].tolist()}
    return out


def cpu_port_run(db, flat_np, offs_np):
    """fallback CPU baseline: the C restatement (oracle/mc_oracle.c), one thread"""
    from oracle import mc_oracle as O
    keys, sizes, values = db.export_part(0)
    tab = O.Table(keys, sizes, values)
    t0 = time.time()
    for i in range(len(offs_np) - 1):
        O.query(tab, flat_np[offs_np[i]:offs_np[i + 1]].tobytes(), b"")
    return time.time() - t0


def host_info():
    """what the host-bound figures ran on: CPU model and usable threads (the e2e step is bound by how fast
    the host cores read ASCII bases)"""
    model = None
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"cpu_model": model, "threads": os.cpu_count()}


def host_sample(flat, offs, n):
    """first n reads as host arrays (bases, offsets starting at 0)"""
    o = offs[:n + 1].cpu().numpy()
    return flat[int(o[0]):int(o[-1])].cpu().numpy(), o - o[0]



def e2e_host_buffers(args, L, db, sk, host_reads, host_offs, nq, top_first, device):
    """End to end through the query_batch seam, from the CALLER'S host buffers to candidates in host
    memory: worker threads (one per host core, three slots each, as the reference's query workers own one
    query_batch host slot each, database_query.hpp:87-124) take the next chunk of reads that is due, add it -
    which packs the bases to 2 bits + ambiguity bit into the slot's pinned buffers -, submit (H2D + kernels +
    D2H on the slot's stream) and wait.  Timed by wall clock over K back-to-back steps, host work included."""
    import threading
    import torch
    from metacache_b200 import _lib
    T = max(1, min((os.cpu_count() or 1) // max(1, args.gpus), args.e2e_threads or 32))    # the ranks of a box share its cores
    per = min(args.slot_reads, (nq + T - 1) // T)
    chunks = [(c, min(c + per, nq)) for c in range(0, nq, per)]
    SPW = max(2, args.e2e_slots_per_worker)                    # slots per worker: fill one while the others are in flight
    nslots = SPW * T
    slot_bases = max(int(host_offs[e] - host_offs[b]) for b, e in chunks)
    qb = _lib.check_ptr(L.mcb200_batch_create(db._h, per, slot_bases + 64, MAXC, 0, nslots))
    tops = np.zeros((nq, MAXC, 4), np.uint32)
    base_ptr = host_reads.ctypes.data
    chunk_offs = [np.ascontiguousarray(host_offs[b:e + 1] - host_offs[b]) for b, e in chunks]
    h2d = sum(((int(host_offs[e] - host_offs[b]) + 31) // 32) * 12 + (e - b + 1) * 4 + (e - b) * 8 for b, e in chunks)
    d2h = nq * MAXC * 16
    errors = []
    spent = [[0.0, 0.0, 0.0] for _ in range(T)]               # per worker: wait + collect | add (pack) | submit

    verify_n = min(len(top_first), nq) if top_first is not None else 0

    def collect(slot, ci):
        """wait: the candidates of the chunk are then in the slot's pinned host buffer (the D2H copy is part of
        the submit); they are copied out only where they are compared with the device-resident path"""
        _lib.check(L.mcb200_batch_wait(qb, slot))
        b, e = chunks[ci]
        if b < verify_n:
            src = L.mcb200_batch_top_candidates(qb, slot, 0)
            C.memmove(tops[b:e].ctypes.data, src, (e - b) * MAXC * 16)

    def worker(t, steps, start, take):
        try:
            pending = [None] * SPW
            k = 0
            start.wait()
            while True:
                # the next chunk of the job, whichever worker is free: K steps x len(chunks) chunks back to back
                # (a static split leaves half the workers idle for a third of the step when the chunks do not
                # divide by the workers: 40 chunks on 16 threads cost 3 rounds instead of 2.5)
                n = take()
                if n >= steps * len(chunks):
                    break
                ci = n % len(chunks)
                j = k % SPW
                slot = SPW * t + j
                t0 = time.perf_counter()
                if pending[j] is not None:
                    collect(slot, pending[j])
                _lib.check(L.mcb200_batch_clear(qb, slot))
                t1 = time.perf_counter()
                b, e = chunks[ci]
                added = _lib.check(L.mcb200_batch_add_reads(qb, slot, base_ptr + int(host_offs[b]),
                                                            chunk_offs[ci].ctypes.data, e - b, 0, 0, SK["winstride"]))
                assert added == e - b
                t2 = time.perf_counter()
                _lib.check(L.mcb200_batch_submit(qb, slot, C.byref(sk)))
                t3 = time.perf_counter()
                spent[t][0] += t1 - t0; spent[t][1] += t2 - t1; spent[t][2] += t3 - t2
                pending[j] = ci
                k += 1
            t0 = time.perf_counter()
            for i in range(SPW):
                j = (k + i) % SPW
                if pending[j] is not None:
                    collect(SPW * t + j, pending[j])
            spent[t][0] += time.perf_counter() - t0
        except Exception as ex:                                   # surfaced by the caller
            errors.append(ex)

    def run(steps):
        for x in spent:
            x[:] = [0.0, 0.0, 0.0]
        start = threading.Barrier(T + 1)
        ticket, lock = [0], threading.Lock()

        def take():
            with lock:
                ticket[0] += 1
                return ticket[0] - 1
        th = [threading.Thread(target=worker, args=(t, steps, start, take)) for t in range(T)]
        for x in th:
            x.start()
        torch.cuda.synchronize(device)
        start.wait()
        t0 = time.perf_counter()
        for x in th:
            x.join()
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        if errors:
            raise errors[0]
        return dt * 1e3

    run(max(args.warmup, 3))
    wall_ms = run(args.steps) / args.steps
    host_ms = {k_: round(1e3 * sum(x[i] for x in spent) / T / args.steps, 3)
               for i, k_ in enumerate(("wait_results", "add_reads_pack", "submit"))}
    same = None
    if top_first is not None:
        n = verify_n
        same = bool(np.array_equal(tops[:n], np.ascontiguousarray(top_first[:n]).view(np.uint32).reshape(n, MAXC, 4)))

    # the same slots already filled (no host packing in the timed region): H2D + kernels + D2H only,
    # device-timed across the slots' streams; a sample of min(#chunks, #slots) chunks
    ns = min(len(chunks), nslots)
    for s_ in range(nslots):
        _lib.check(L.mcb200_batch_clear(qb, s_))
    for s_ in range(ns):
        b, e = chunks[s_]
        _lib.check(L.mcb200_batch_add_reads(qb, s_, base_ptr + int(host_offs[b]), chunk_offs[s_].ctypes.data,
                                            e - b, 0, 0, SK["winstride"]))
    pre = []
    for it in range(3 + args.steps):
        for s_ in range(ns):
            _lib.check(L.mcb200_batch_submit(qb, s_, C.byref(sk)))
        for s_ in range(ns):
            _lib.check(L.mcb200_batch_wait(qb, s_))
        span = C.c_float(0)
        _lib.check(L.mcb200_batch_span_ms(qb, 0, ns, C.byref(span)))
        if it >= 3:
            pre.append(span.value)
    pre_reads = sum(chunks[s_][1] - chunks[s_][0] for s_ in range(ns))
    pre_ms = sum(pre) / len(pre) * (nq / pre_reads)
    L.mcb200_batch_destroy(qb)
    return {"value": nq / (wall_ms * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": round(wall_ms, 3), "timed_by": "wall clock, host work included",
            "host_threads": T, "slots": nslots, "slots_per_worker": SPW, "reads_per_slot": per, "host_ms_per_thread_per_step": host_ms,
            "chunks": "taken one by one by whichever worker is free",
            "packer": ("scalar", "avx2", "avx512")[int(L.mcb200_internal_pack_has_avx2())],
            "results_equal_device_resident_path": same,
            "prefilled": {"value": nq / (pre_ms * 1e-3), "ms_per_step": round(pre_ms, 3),
                          "note": "slots filled (packed) before the timed region: pinned -> H2D -> kernels -> D2H only, "
                                  "CUDA events across the slots' streams, scaled from %d reads" % pre_reads},
            "api": "mcb200_batch_add_reads (packs 2 bit/base on the host) + submit + wait, from the caller's "
                   "ASCII buffers to candidates in host memory (query_batch seam)"}


def compare_target_sharded(args, rank, world, device, stream, nq, n_bases, SK_, maxc, dist, barrier):
    """The reference's partitioning on the same reads, in the same run: rank r holds part r, NCCL all-gather of
    sketches, every GPU probes every read, all-to-all of partial top hits, on-device merge.
    Short (2 warm-up + 3 timed steps); device-resident; returned for the bench line next to the merged mode."""
    import torch
    from metacache_b200 import _lib
    from metacache_b200._lib import DevQueries, Sketching
    from metacache_b200.distributed import ShardedQuery
    L = _lib.lib()
    # everything that can fail on one rank alone (allocations) happens before the first collective, and the
    # ranks agree to go on: a rank that gave up must not leave the others waiting in an all-gather
    err = None
    try:
        db, bases, _, info = build_part(args, rank, device)
        flat, offs = make_reads(args, bases, rank, device)
        del bases
        sk = Sketching(**SK_)
        seq_off = offs.to(torch.int32)
        seq_qry = torch.arange(nq, dtype=torch.int32, device=device)
        max_win = (2 + (offs[1:] - offs[:-1]) // SK_["winstride"]).to(torch.int32)
        ws = _lib.check_ptr(L.mcb200_workspace_create(db._h, nq, nq, n_bases + 64, maxc, 0))
        q = DevQueries(flat.data_ptr(), seq_off.data_ptr(), seq_qry.data_ptr(), max_win.data_ptr(), nq, nq, n_bases)
        sq = ShardedQuery(db, ws, nq, 2 * nq, SK_["sketchlen"], maxc, device, stream)
        torch.cuda.synchronize(device)
    except Exception as ex:                                  # noqa: BLE001
        err = ex
    flag = torch.tensor([int(err is None)], dtype=torch.int64, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not int(flag.item()):
        return {"value": None, "error": f"setup failed on a rank: {type(err).__name__ if err else 'other rank'}: {str(err)[:160] if err else ''}"}
    try:
        for _ in range(2):
            sq.step(q, sk, max_win)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        steps = 3
        for _ in range(steps):
            sq.step(q, sk, max_win)
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / steps
        _lib.check(L.mcb200_workspace_check(ws))
        L.mcb200_workspace_destroy(ws)
        db.close()
        return {"value": nq * world / (ms * 1e-3), "unit": "reads/s", "ms_per_step": round(ms, 3), "steps": steps,
                "what": "one part per GPU, every GPU probes every read, NCCL all-gather of sketches + all-to-all of "
                        "partial top hits + on-device merge (--shard-by target), device-resident"}
    except Exception as ex:                                  # never lose the main numbers
        return {"value": None, "error": f"{type(ex).__name__}: {str(ex)[:200]}"}


def choose_shard_mode(args, world, device_bytes):
    """How the N-part database is held at N > 1: "single" (one GPU / replicas of one part),
    "merged" (all parts in one table on every GPU, reads sharded), "target" or "feature" (database sharded)."""
    mode = args.shard_by
    if args.replicate_merged:
        mode = "merged"
    if args.merged_parts > 1:
        return "merged"
    if world == 1 or args.replicate:
        return "single"
    if mode == "auto":
        # the sharded load keeps every part's (feature, location) pairs twice for a moment (collected + merged,
        # 8 B per location) plus sort keys and the final table: ~26 B per location at the peak
        need = world * (args.targets * (args.target_len / SK["winstride"]) * SK["sketchlen"]) * 26.0
        mode = "merged" if need < 0.85 * device_bytes else "target"
    return mode


def reference_sample(args, nq, threads, per_read_scale):
    return args.cpu_sample or int(min(nq, max(200_000, 150_000 * threads) * per_read_scale))


def reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the hot path on this box's host
    cores.  This process never loads libmcb200.so or torch: the synthetic database (in the reference's
    on-disk format) and the read sample are produced by a CHILD process (`--prepare-reference`, untimed;
    building DB-S with the reference's own `metacache build` would take hours) unless they are cached."""
    threads = os.cpu_count() or 1
    base = os.path.join(db_cache_dir(args, 0), "db")
    prepared = base + f".{args.workload}.prepared.json"
    if not (os.path.exists(prepared) and os.path.exists(base + ".ok")):
        env = {k: v for k, v in os.environ.items()
               if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "LOCAL_WORLD_SIZE",
                            "GROUP_RANK", "ROLE_RANK", "TORCHELASTIC_RUN_ID")}
        cmd = [sys.executable, os.path.abspath(__file__), "--prepare-reference", "--workload", args.workload,
               "--targets", str(args.targets), "--target-len", str(args.target_len), "--cache", args.cache,
               "--reads", str(args.reads), "--cpu-sample", str(args.cpu_sample)]
        subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL)
    info = json.load(open(prepared))
    config, sample = info["config"], info["sample"]
    if args.gpus > 1:
        config["reference_db"] = "part 0 only (1/%d of the sharded database)" % args.gpus
    from oracle import refio
    passes = args.warmup + args.steps
    if os.path.exists(refio.HARNESS):
        r = refio.run_harness(info["base"], info["reads"], "-", threads=threads, repeat=passes, sketches=0,
                              allhits=0, maxcand=MAXC)
        tt = r["passes"][-args.steps:]
        val, kind, cores, ms = sample * len(tt) / sum(tt), "reference", threads, 1e3 * sum(tt) / len(tt)
        sample_desc = (f"first {sample} reads of the workload per step, reference hot path "
                       f"(database::query_host) via oracle/_ref/mc_ref_harness, {threads} threads, "
                       f"db load {r['load_seconds']:.0f}s untimed")
    else:
        from metacache_b200 import dbformat
        from oracle import mc_oracle as O
        part = dbformat.read_cache(info["base"] + ".cache0")
        tab = O.Table(part.keys, part.sizes, part.values)
        lines = open(info["reads"], "rb").read().split(b"\n")[:20000]
        t0 = time.time()
        for ln in lines:
            O.query(tab, ln, b"")
        t = time.time() - t0
        val, kind, cores, ms = len(lines) / t, "port", 1, t * 1e3
        sample_desc = f"first {len(lines)} reads of the workload, oracle/mc_oracle.c, 1 thread"
    print(json.dumps({"impl": "reference", "metric": info["metric"], "value": val, "unit": "reads/s",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/u64",
                      "data": "synthetic", "config": config,
                      "cpu_baseline": {"value": val, "unit": "reads/s", "cores": cores, "kind": kind,
                                       "sample": sample_desc, "host": host_info()},
                      "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


# --------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if args.impl == "reference":
        return 0 if rank != 0 else reference_arm(args)
    if args.prepare_reference:
        world = 1

    import torch
    from metacache_b200 import _lib
    from metacache_b200._lib import DevQueries, Sketching

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    device = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(device)
    dist = None
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    L = _lib.lib()
    mode = choose_shard_mode(args, world, torch.cuda.get_device_properties(device).total_memory)
    if args.workload == "C3" and mode in ("target", "feature"):
        raise SystemExit("workload C3 with a sharded database: use --shard-by merged or --replicate for N > 1")
    args.replicate_merged = mode == "merged"
    args.shard_by = mode if mode in ("target", "feature") else "target"
    sharded = mode in ("target", "feature")                  # else: every GPU answers its own reads alone
    part = rank if sharded else 0
    threads = os.cpu_count() or 1

    if not args.reads:
        args.reads = 10_000_000 if args.workload == "C2" else 2_000_000
    by_feature = sharded and args.shard_by == "feature"
    if by_feature:
        db, bases, wins, dbinfo = build_feature_shard(args, rank, world, device)
    elif args.replicate_merged and world == 1:
        db, bases, wins, dbinfo = build_feature_shard(args, rank, world, device, all_features=True)
    elif args.replicate_merged:
        # if the merged table does not fit on some rank, every rank falls back to the target-sharded database
        try:
            db, bases, wins, dbinfo = build_feature_shard(args, rank, world, device, all_features=True)
            ok = 1
        except (RuntimeError, MemoryError) as ex:
            sys.stderr.write(f"[rank {rank}] merged table not built ({type(ex).__name__}: {str(ex)[:200]}): target-sharded instead\n")
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int64, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not int(flag.item()):
            db = bases = wins = dbinfo = None
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            args.replicate_merged, sharded, part = False, True, rank
            db, bases, wins, dbinfo = build_part(args, part, device)
    else:
        db, bases, wins, dbinfo = build_part(args, part, device)
    flat, offs = make_reads(args, bases, rank, device)       # uint8 bases back to back + int64 offsets, on the device
    del bases
    torch.cuda.empty_cache()
    nq = args.reads
    n_bases = int(offs[-1].item())
    part_desc = ("single partition" if world == 1 else
                 f"{world}-way target-partitioned database ({world * args.targets} targets) sharded by FEATURE over "
                 f"{world} GPUs, features and location lists exchanged over NCCL" if by_feature else
                 f"{world}-way target-partitioned, one part per GPU, every GPU probes every read" if sharded else
                 f"{world}-way target-partitioned database ({world * args.targets} targets), all parts merged into one "
                 f"table on every GPU, every GPU queries its own reads (no exchange in the data path)" if args.replicate_merged else
                 f"single partition replicated on {world} GPUs, every GPU queries its own reads")
    if args.workload == "C2":
        metric = "reads_per_second_150bp"
        workload = (f"C2: {nq} x {READ_LEN}bp synthetic reads (R150) vs {args.targets}-target x "
                    f"{args.target_len}bp synthetic db (DB-S, k=16 s=16 w=127), " + part_desc)
        read_len = READ_LEN
    else:
        metric = "reads_per_second_long"
        workload = (f"C3: {nq} synthetic long reads (RLONG: 200-19000 bp, median 480, {n_bases / nq:.0f} bp mean) vs "
                    f"{args.targets}-target x {args.target_len}bp synthetic db (DB-S, k=16 s=16 w=127), " + part_desc)
        read_len = round(n_bases / nq, 1)
    config = {"workload": workload, "reads_per_gpu": nq, "read_len": read_len, "targets_per_part": args.targets,
              "db": dbinfo, "l2": "inputs larger than L2 (reads %.1f GB + table %.1f GB per step)" %
              (n_bases / 1e9, dbinfo["table_gb"]), "parallelism": ("feature-sharded x%d" if by_feature else "db-sharded x%d" if sharded or world == 1 else
                              "merged replicas x%d" if args.replicate_merged else "replicas x%d") % world}
    per_read_scale = READ_LEN * nq / n_bases                 # CPU samples are sized in 150 bp read equivalents

    if args.prepare_reference:
        # child of the reference arm: database of part 0 in the reference's format + the sample as text
        base = export_reference_db(args, db, wins, 0)
        sample = reference_sample(args, nq, threads, per_read_scale)
        rt = reference_reads_path(args, base, sample)
        if not os.path.exists(rt):
            flat_np, offs_np = host_sample(flat, offs, sample)
            write_reads_txt(rt + ".tmp", flat_np, offs_np)
            os.replace(rt + ".tmp", rt)
        json.dump({"base": base, "reads": rt, "sample": sample, "n_bases": n_bases, "config": config,
                   "metric": metric}, open(base + f".{args.workload}.prepared.json", "w"))
        return 0

    # ---------------- device-resident inputs ----------------
    sk = Sketching(**SK)
    if n_bases >= (1 << 32) - 4096:
        raise SystemExit("batch too large for 32-bit base offsets: lower --reads")
    seq_off = offs.to(torch.int32)                            # (values < 2^32 wrap into the u32 the library reads)
    seq_qry = torch.arange(nq, dtype=torch.int32, device=device)
    max_win = (2 + (offs[1:] - offs[:-1]) // SK["winstride"]).to(torch.int32)
    stream = torch.cuda.Stream(device)
    sp = C.c_void_p(stream.cuda_stream)
    nq_total = nq * world
    ws = _lib.check_ptr(L.mcb200_workspace_create(db._h, nq, nq, n_bases + 64, MAXC, 0))
    if args.table_slots:
        _lib.check(L.mcb200_workspace_set_warp_capacity(ws, args.table_slots))
    q = DevQueries(flat.data_ptr(), seq_off.data_ptr(), seq_qry.data_ptr(), max_win.data_ptr(), nq, nq, n_bases)
    d_top = torch.empty((nq, MAXC, 4), dtype=torch.int32, device=device)

    if not sharded:
        def step():
            _lib.check(L.mcb200_query_device(ws, C.byref(q), C.byref(sk), d_top.data_ptr(), sp))
    elif by_feature:
        from metacache_b200.distributed import DeviceBackend, FeatureShardedQuery, TorchComm, feature_sharded_step
        chunk = min(args.chunk_reads, nq)
        fstreams = [torch.cuda.Stream(device) for _ in range(max(1, min(3, args.shard_streams)))]
        fstreams = [fstreams[i % len(fstreams)] for i in range(3)]
        backend = DeviceBackend(db, world, chunk, MAXC, device)
        fq = FeatureShardedQuery(backend, TorchComm(), SK["sketchlen"], MAXC, chunk_queries=chunk,
                                 n_slots=3, streams=fstreams)

        def step():
            with torch.cuda.stream(stream):
                feature_sharded_step(fq, ws, q, sk, max_win, d_top)
    else:
        from metacache_b200.distributed import ShardedQuery
        nwin = 2 * nq                            # 150 bp reads: two windows each
        sq = ShardedQuery(db, ws, nq, nwin, SK["sketchlen"], MAXC, device, stream)
        d_top = sq.top

        def step():
            sq.step(q, sk, max_win)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---------------- value: inputs resident in HBM ----------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    nwin_launch = L.mcb200_workspace_num_windows(ws) if (not sharded or by_feature) else 2 * nq
    if by_feature:
        fq.enable_timing(True)
        fq.stats.update(features_sent=0, locations_received=0, chunks=0)
    cnt = (C.c_uint64 * 8)()
    _lib.check(L.mcb200_workspace_counters(ws, cnt))         # resets the counters
    _lib.check(L.mcb200_workspace_set_profiling(ws, 1))
    launches0 = L.mcb200_kernel_launches()
    clocks = ClockSampler(device.index)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = L.mcb200_kernel_launches() - launches0
    phases = None
    if by_feature:
        phases = {k_: v / args.steps for k_, v in fq.phase_ms().items()}
        fq.enable_timing(False)
    stage = (C.c_float * 8)()
    _lib.check(L.mcb200_workspace_stage_times(ws, stage))
    _lib.check(L.mcb200_workspace_set_profiling(ws, 0))
    _lib.check(L.mcb200_workspace_counters(ws, cnt))
    _lib.check(L.mcb200_workspace_check(ws))                 # no read may have been dropped (scratch overflow)
    # rows of the first reads, kept for the bit-exact comparison with the reference below
    keep = int(min(nq, max(200_000, 100_000 * threads) * per_read_scale)) if not args.cpu_sample else args.cpu_sample
    top_first = d_top[:keep].cpu().numpy() if not sharded else None
    if by_feature and os.environ.get("MCB200_BENCH_DUMP_TOPS"):
        np.save(os.environ["MCB200_BENCH_DUMP_TOPS"] + f".r{rank}.npy", d_top[:200_000].cpu().numpy())
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = nq_total / (ms_step * 1e-3)

    # roofline of the dominant kernel (fused probe + sort + candidates)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    calls = float(args.steps)
    n_launch = calls * (world if sharded else 1)               # fused-kernel launches of this rank in the timed region
    feats_probed, locs, sectors = cnt[4] / n_launch, cnt[3] / n_launch, cnt[5] / n_launch
    list_lines = cnt[6] / n_launch
    alg_bytes = 4 * SK["sketchlen"] * nwin_launch + 16 * feats_probed + 8 * locs + 16 * MAXC * nq + 8 * nq
    k_ms = (stage[3] + stage[4]) / n_launch
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    # DRAM bytes per read of the dominant kernel from the committed `ncu --set full` capture
    # (profiles/traffic.json; measured on a 1 M-read launch of the same workload), scaled to this launch
    traffic = None
    try:
        cal = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if not sharded:
            traffic = int(cal["query_fast_kernel"]["dram_bytes_per_read"] * nq)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "query_fast_kernel (+ query_heavy_kernel): fused probe / aggregate / top-hits",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                "alg_bytes_per_launch": int(alg_bytes), "kernel_ms_per_launch": round(k_ms, 3),
                "per_read": {"features": round(feats_probed / nq, 2), "locations": round(locs / nq, 2),
                             "table_sectors_32B": round(sectors / nq, 2)},
                "stage_ms_per_step": {"encode": round(stage[0] / calls, 3), "window_tables": round(stage[1] / calls, 3),
                                      "sketch": round(stage[2] / calls, 3), "probe_reduce_warp": round(stage[3] / calls, 3),
                                      "probe_reduce_heavy": round(stage[4] / calls, 3), "merge": round(stage[5] / calls, 3)},
                "sketch_kernel": {"alg_bytes": int(0.375 * n_bases + 4 * SK["sketchlen"] * nwin_launch),
                                  "achieved": round((0.375 * n_bases + 64 * nwin_launch) / max(stage[2] / calls, 1e-6) / 1e6, 1),
                                  "unit": "GB/s"},
                # what the memory system allows for this access pattern: every table bucket and every 64-byte line of a
                # location list is an isolated access = one HBM line activation; profiles/gather_bench_r1.log measures
                # how many of those the B200 serves per second (independent 32-byte requests over a 16 GB working set)
                "random_access": (lambda lines, rate: {
                    "lines_per_read": round(lines / nq, 2), "measured_lines_per_s": rate,
                    "floor_ms_per_launch": round(lines / rate * 1e3, 3),
                    "frac_of_floor": round((lines / rate * 1e3) / k_ms, 4) if k_ms > 0 else None,
                    "source": "profiles/gather_bench_r1.log (granule 32 B, 16 GB working set)"})(sectors + list_lines, 43.45e9),
                "queries_fused_warp": int(cnt[0] / n_launch), "queries_cta_smem": int(cnt[1] / n_launch),
                "queries_cta_global": int(cnt[2] / n_launch)}

    if by_feature:
        f_sent = fq.stats["features_sent"] / calls
        l_recv = fq.stats["locations_received"] / calls
        p_ms = phases["probe"]
        pb = f_sent * 32.0                                       # 4 B feature in, 16 B slot, 4 + 8 B out (per rank ~ balanced)
        roofline = {"bound": "hbm", "kernel": "shard_probe_kernel (+ scan): owner-side slot lookup, one thread per feature",
                    "achieved": round(pb / (p_ms * 1e-3) / 1e9, 1) if p_ms > 0 else 0.0, "peak": peak, "unit": "GB/s",
                    "frac": round(pb / (p_ms * 1e-3) / 1e9 / peak, 4) if p_ms > 0 else 0.0, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "alg_bytes_per_launch": int(pb), "kernel_ms_per_launch": round(p_ms, 3),
                    "note": "rank 0; CUDA events around each operation on its chunk's stream, summed over the chunks of a step; "
                            "operations of different chunks overlap on the GPU, so the sum exceeds the step",
                    "phase_ms_per_step": {k_: round(v, 3) for k_, v in phases.items()},
                    "stage_ms_per_step": {"encode": round(stage[0] / calls, 3), "window_tables": round(stage[1] / calls, 3),
                                          "sketch": round(stage[2] / calls, 3)},
                    "per_read": {"features_routed": round(f_sent / nq, 2), "locations_returned": round(l_recv / nq, 2)},
                    "exchange_bytes_per_step_per_rank": int(f_sent * 8 + l_recv * backend.loc_bytes),
                    "chunks_per_step": int(fq.stats["chunks"] / calls)}

    # ---------------- e2e: host buffers through the batch API (H2D + kernels + D2H) -----------
    e2e = None
    if args.no_e2e:
        host_reads = host_offs = None
    elif not sharded:
        host_reads = flat.cpu().numpy()
        host_offs = offs.cpu().numpy().astype(np.uint64)
        del d_top
        L.mcb200_workspace_destroy(ws)
        ws = None
        del flat
        torch.cuda.empty_cache()
        e2e = e2e_host_buffers(args, L, db, sk, host_reads, host_offs, nq, top_first, device)
        if dist is not None:                                     # replicas: the slowest rank sets the step
            t = torch.tensor([e2e["ms_per_step"], e2e["prefilled"]["ms_per_step"]], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e["ms_per_step"] = round(float(t[0].item()), 3)
            e2e["value"] = nq_total / (e2e["ms_per_step"] * 1e-3)
            e2e["prefilled"]["ms_per_step"] = round(float(t[1].item()), 3)
            e2e["prefilled"]["value"] = nq_total / (float(t[1].item()) * 1e-3)
            e2e["h2d_bytes_per_step"] *= world
            e2e["d2h_bytes_per_step"] *= world
    else:
        # pinned host slice -> device, distributed pipeline, final tops of my slice -> pinned host
        pin_in = torch.empty(n_bases, dtype=torch.uint8).pin_memory()
        pin_in.copy_(flat.cpu())
        pin_out = torch.empty((nq, MAXC, 4), dtype=torch.int32).pin_memory()

        def e2e_step():
            with torch.cuda.stream(stream):
                flat.copy_(pin_in, non_blocking=True)
            step()
            with torch.cuda.stream(stream):
                pin_out.copy_(d_top, non_blocking=True)

        for _ in range(3):
            e2e_step()
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            e2e_step()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item()) / args.steps
        e2e = {"value": nq_total / (e2e_ms * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": int(n_bases * world),
               "d2h_bytes_per_step": int(nq * MAXC * 16 * world), "ms_per_step": round(e2e_ms, 3),
               "api": "pinned host reads -> H2D -> " + ("sketch / route / all-to-all / probe / gather / all-to-all / reduce"
                                                          if by_feature else "sketch/all-gather/probe/all-to-all/merge") + " -> D2H"}

    clk = clocks.stop()                                      # sampled through both timed regions (value and e2e)

    # ---------------- merged mode: the target-sharded step measured beside it, same run, same reads ---------
    target_cmp = None
    if args.replicate_merged and world > 1 and dist is not None and not args.no_compare_target and args.workload == "C2":
        target_cmp = compare_target_sharded(args, rank, world, device, stream, nq, n_bases, SK, MAXC, dist, barrier)

    # ---------------- CPU baseline beside it (rank 0, N = 1 only) ----------------
    cpu = None
    parity = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline and not args.no_e2e:
        try:
            base = export_reference_db(args, db, wins, 0)
            sample = args.cpu_sample or int(min(nq, max(200_000, 100_000 * threads) * per_read_scale))
            offs_np = host_offs[:sample + 1].astype(np.int64)
            flat_np = host_reads[:int(offs_np[-1])]
            tops_file = base + f".tops{sample}.bin"
            r = cpu_reference_run(base, flat_np, offs_np, threads, 2, tops=tops_file)
            if r is not None:
                parity = parity_against_reference(tops_file, top_first[:sample])
                os.unlink(tops_file)
                cpu = {"value": sample / r["passes"][-1], "unit": "reads/s", "cores": threads, "kind": "reference",
                       "sample": f"first {sample} reads of the workload, reference hot path (database::query_host) "
                                 f"via oracle/_ref/mc_ref_harness, {threads} threads, 2nd of 2 passes; "
                                 f"db load {r['load_seconds']:.0f}s untimed"}
            else:
                n = int(20000 * per_read_scale)
                tsec = cpu_port_run(db, flat_np, offs_np[:n + 1])
                cpu = {"value": n / tsec, "unit": "reads/s", "cores": 1, "kind": "port",
                       "sample": f"first {n} reads of the workload, oracle/mc_oracle.c, 1 thread"}
        except Exception as ex:                                  # never lose the GPU numbers
            cpu = {"value": None, "unit": "reads/s", "cores": threads, "kind": "reference",
                   "sample": f"failed: {type(ex).__name__}: {ex}"}

    if rank == 0:
        out = {"metric": metric, "value": value, "unit": "reads/s", "n_gpus": world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u32/u64", "data": "synthetic", "config": config,
               "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity}
        out["host"] = host_info()
        if target_cmp is not None:
            out["target_sharded_same_run"] = target_cmp
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
