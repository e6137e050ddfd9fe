#!/usr/bin/env python
"""Generates tests/golden/reader.json: FASTA/FASTQ edge-case files and what the UNMODIFIED reference
reader (oracle/_ref/mc_ref_reader = ref_reader.cpp + the reference's sequence_io.cpp) returns for them.
Runs only where /root/reference is present (this container).  Test infrastructure only."""
import base64
import gzip
import json
import os
import subprocess
import sys
import tempfile
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_READER = os.path.join(ROOT, "oracle", "_ref", "mc_ref_reader")


def big_content(recipe):
    """deterministic large inputs (not stored in the fixture): see tests/test_reader.py"""
    sys.path.insert(0, ROOT)
    from tests.reader_cases import big_content as bc
    return bc(recipe)


def run_ref(paths):
    out = subprocess.run([REF_READER] + paths, capture_output=True).stdout
    recs = []
    for line in out.split(b"\n"):
        if not line:
            continue
        f = line.split(b"\t")
        if f[0] == b"ERROR":
            return {"error": f[1].decode()}
        recs.append([int(f[0]), f[1].decode("latin1"), f[2].decode("latin1"), f[3].decode("latin1")])
    return {"records": recs}


def main():
    sys.path.insert(0, ROOT)
    from tests.reader_cases import CASES
    out = []
    with tempfile.TemporaryDirectory() as d:
        for c in CASES:
            paths = []
            for i, (kind, data) in enumerate(c["files"]):
                raw = big_content(data) if kind == "big" else data
                if c.get("gz"):
                    raw = gzip.compress(raw)
                p = os.path.join(d, f"{c['name']}_{i}" + (".gz" if c.get("gz") else ""))
                open(p, "wb").write(raw)
                paths.append(p)
            if c.get("pairseq"):
                paths = [paths[0], paths[0]]
            res = run_ref(paths)
            entry = {"name": c["name"]}
            if "error" in res:
                entry["error"] = res["error"]
            elif any(k == "big" for k, _ in c["files"]):
                recs = res["records"]
                entry["digest"] = {"n": len(recs),
                                   "crc": zlib.crc32(json.dumps(recs, separators=(",", ":")).encode()) & 0xFFFFFFFF,
                                   "lens": [len(r[2]) for r in recs[:50]]}
            else:
                entry["records"] = res["records"]
            out.append(entry)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "reader.json"), "w"), indent=0)
    print(f"{len(out)} cases written")


if __name__ == "__main__":
    main()
