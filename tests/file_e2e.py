"""Measurement script (not a pytest module; it lives under tests/ because it times the reference CLI from
oracle/_ref beside our path, which only tests/ and bench.py may do).

File in -> top hits out: `query_file` (reader threads + batch slots) on a FASTA file of R150 reads in
/dev/shm vs the reference CLI (`metacache query`, all host threads) on the same file and database."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from metacache_b200.reader import query_file, SequenceReader  # noqa: E402

NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
REF_SAMPLE = 2_000_000


class A:
    targets, target_len, load_factor, reads, workload = 50_000, 100_000, 0.0, NQ, "C2"
    replicate_merged, merged_parts = False, 0
    cache = "/dev/shm/mcb200_bench"


dev = torch.device("cuda", 0)
db, bases, wins, info = bench.build_part(A, 0, dev)
flat, offs = bench.make_reads(A, bases, 0, dev)
del bases
reads = flat.cpu().numpy().reshape(NQ, 150)
path = "/dev/shm/mcb200_r150.fa"
t0 = time.time()
with open(path, "wb") as f:
    pw = 10 ** np.arange(8, -1, -1)
    for i in range(0, NQ, 500_000):
        blk = reads[i:i + 500_000]
        n = len(blk)
        rec = np.empty((n, 163), np.uint8)                     # ">r%09d\n" + 150 bases + "\n"
        rec[:, 0], rec[:, 1], rec[:, 11], rec[:, 162] = ord(">"), ord("r"), 10, 10
        rec[:, 2:11] = (np.arange(i, i + n)[:, None] // pw[None, :]) % 10 + 48
        rec[:, 12:162] = blk
        f.write(rec.tobytes())
size = os.path.getsize(path)
out = {"file": path, "reads": NQ, "file_gb": round(size / 1e9, 3), "write_s": round(time.time() - t0, 1)}
threads = os.cpu_count() or 1
# parse only
for T in (1, min(threads, 16)):
    cuts = [size * i // T for i in range(T + 1)]
    import threading
    rs = [SequenceReader(path, byte_range=(cuts[i], cuts[i + 1])) for i in range(T)]
    res = [None] * T
    def w(i): res[i] = rs[i].skip(1 << 62)
    t0 = time.time(); th = [threading.Thread(target=w, args=(i,)) for i in range(T)]
    [x.start() for x in th]; [x.join() for x in th]; dt = time.time() - t0
    assert sum(r[0] for r in res) == NQ
    out[f"parse_only_{T}_threads_Mreads_s"] = round(NQ / dt / 1e6, 1)
for T in sorted(set([1, 4, min(threads, 8), min(threads, 16)])):
    best = None
    for rep in range(3):
        tm = {}
        t0 = time.time()
        top, _ = query_file(db, path, threads=T, batch_queries=1 << 19, timing=tm)
        tm["wall_s"] = time.time() - t0
        assert len(top) == NQ
        if best is None or tm["run_s"] < best["run_s"]:
            best = tm
    out[f"query_file_{T}_threads"] = {"Mreads_s": round(NQ / best["run_s"] / 1e6, 2), "run_s": round(best["run_s"], 3),
                                      "setup_s": round(best["setup_s"], 3), "wall_s": round(best["wall_s"], 3),
                                      "note": "best of 3; run = parse + H2D + kernels + D2H in all threads, setup = pinned/device buffers"}
# same answers as the device-resident path on a sample
out["mapped_frac"] = round(float((top[:, 0, 1] >= 5).mean()), 4)
# reference CLI on the first REF_SAMPLE reads
base = bench.export_reference_db(A, db, wins, 0)
ref_bin = os.path.join(ROOT, "oracle", "_ref", "metacache")
if os.path.exists(ref_bin):
    sample = "/dev/shm/mcb200_r150_sample.fa"
    with open(path, "rb") as f, open(sample, "wb") as g:
        g.write(f.read(REF_SAMPLE * 163))
    t0 = time.time()
    p = subprocess.run([ref_bin, "query", base, sample, "-threads", str(threads), "-no-map"],
                       capture_output=True, text=True)
    dt = time.time() - t0
    lines = [l.strip() for l in p.stdout.replace("\r", "\n").splitlines()
             if l.startswith("# queries") or l.startswith("# time") or l.startswith("# speed")]
    out["reference_cli"] = {"threads": threads, "reads": REF_SAMPLE, "wall_s_incl_db_load": round(dt, 1), "reported": lines}
    # the same CLI built on libmcb200 (oracle/_ref/metacache_mcb200: the reference's host program, unchanged,
    # with the two seam headers replaced - INTEGRATION.md 1): same sample, then the whole file
    dropin = os.path.join(ROOT, "oracle", "_ref", "metacache_mcb200")
    if os.path.exists(dropin):
        for name, fpath, nreads in (("dropin_cli_sample", sample, REF_SAMPLE), ("dropin_cli_full_file", path, NQ)):
            t0 = time.time()
            p = subprocess.run([dropin, "query", base, fpath, "-threads", str(threads), "-no-map"],
                               capture_output=True, text=True)
            dt = time.time() - t0
            lines = [l.strip() for l in p.stdout.replace("\r", "\n").splitlines()
                     if l.startswith("# queries") or l.startswith("# time") or l.startswith("# speed")]
            out[name] = {"threads": threads, "reads": nreads, "wall_s_incl_db_load": round(dt, 1), "reported": lines,
                         "rc": p.returncode, "stderr_tail": p.stderr[-200:] if p.returncode else ""}
    os.unlink(sample)
os.unlink(path)
print(json.dumps(out))
