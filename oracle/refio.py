"""Helpers around the reference harness output - TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
HARNESS = os.path.join(REF_DIR, "mc_ref_harness")
METACACHE = os.path.join(REF_DIR, "metacache")


def read_fasta(path):
    """-> list of (header, sequence bytes).  Plain FASTA/FASTQ, enough for test data."""
    out, hdr, seq = [], None, []
    with open(path, "rb") as f:
        first = f.read(1)
        f.seek(0)
        if first == b"@":
            lines = f.read().split(b"\n")
            for i in range(0, len(lines) - 3, 4):
                out.append((lines[i][1:].decode(), lines[i + 1].strip()))
            return out
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if hdr is not None:
                    out.append((hdr, b"".join(seq)))
                hdr, seq = line[1:].decode(), []
            else:
                seq.append(line)
        if hdr is not None:
            out.append((hdr, b"".join(seq)))
    return out


def write_reads_txt(path, reads):
    """reads: list of bytes (single) or (bytes, bytes) pairs."""
    with open(path, "wb") as f:
        for r in reads:
            if isinstance(r, (tuple, list)):
                f.write(r[0] + b" " + r[1] + b"\n")
            else:
                f.write(r + b"\n")


def parse_harness_output(path):
    """-> list of dicts {sketches: [np.uint32...], allhits: np.uint64 (tgt<<32|win), top: [(tgt,hits,beg,end)]}"""
    a = np.fromfile(path, dtype="<u4")
    assert a[0] == 0x4D435246, "bad magic"
    n = int(a[1])
    p = 2
    res = []
    for _ in range(n):
        nsk = int(a[p]); p += 1
        sks = []
        for _ in range(nsk):
            ln = int(a[p]); p += 1
            sks.append(a[p:p + ln].copy()); p += ln
        nall = int(a[p]); p += 1
        wt = a[p:p + 2 * nall].reshape(nall, 2); p += 2 * nall
        allh = (wt[:, 1].astype(np.uint64) << np.uint64(32)) | wt[:, 0].astype(np.uint64)
        ntop = int(a[p]); p += 1
        top = [tuple(int(x) for x in a[p + 4 * i:p + 4 * i + 4]) for i in range(ntop)]
        p += 4 * ntop
        res.append({"sketches": sks, "allhits": allh, "top": top})
    assert p == len(a)
    return res


def run_harness(db, reads_txt, out_bin="-", **kw):
    """-> (stdout stats dict).  kw: maxcand, insert, part, threads, repeat, sketches, allhits"""
    cmd = [HARNESS, db, reads_txt, out_bin] + [f"{k}={v}" for k, v in kw.items()]
    lines = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout.strip().splitlines()
    out = {k: float(v) for k, v in (kv.split("=") for kv in lines[-1].split())}
    out["passes"] = [float(l.split("seconds=")[1]) for l in lines[:-1] if l.startswith("pass=")]
    return out
