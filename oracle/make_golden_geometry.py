"""tests/golden/g3.npz: reference outputs for NON-default sketching geometries - TEST INFRASTRUCTURE ONLY.

For every geometry: a small database built by the reference's `metacache build -kmerlen K -sketchlen S
-winlen W -winstride L` from the g1 targets, and the reference's sketches / all-hits / top candidates
(oracle/ref_harness.cpp) for ~230 of the g1 reads (all edge cases included).  Runs only where
/root/reference is present.   usage: python oracle/make_golden_geometry.py"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refio  # noqa: E402
from oracle.make_golden import harness_to_arrays  # noqa: E402
from metacache_b200 import dbformat  # noqa: E402
from tests.golden_util import G1  # noqa: E402

GEOMETRIES = [(12, 8, 60, 49), (16, 16, 256, 241), (16, 32, 127, 112), (9, 16, 300, 292), (16, 16, 40, 25)]


def main():
    g = G1()
    tmp = tempfile.mkdtemp(prefix="mcgold3_")
    names = [f"T{i}" for i in range(len(g.targets))]
    fa = os.path.join(tmp, "t.fa")
    with open(fa, "wb") as f:
        for n, s in zip(names, g.targets):
            f.write(b">" + n.encode() + b"\n" + s + b"\n")
    # every 4th read + the last 60 single reads (edge cases) + every 5th pair
    singles = [i for i, (a, b) in enumerate(g.reads) if not b]
    pairs = [i for i, (a, b) in enumerate(g.reads) if b]
    pick = sorted(set(singles[::4] + singles[-60:] + pairs[::5]))
    reads = [g.reads[i] for i in pick]

    def norm(x):
        return x if len(x) else b"-"
    rt = os.path.join(tmp, "reads.txt")
    refio.write_reads_txt(rt, [norm(a) if not b else (norm(a), norm(b)) for a, b in reads])
    out = {"read_index": np.asarray(pick, np.int32), "geometries": np.asarray(GEOMETRIES, np.uint32)}
    for gi, (k, s, w, st) in enumerate(GEOMETRIES):
        db = os.path.join(tmp, f"db{gi}")
        subprocess.check_call([refio.METACACHE, "build", db, fa, "-parts", "1", "-silent", "-kmerlen", str(k),
                               "-sketchlen", str(s), "-winlen", str(w), "-winstride", str(st)],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        meta = dbformat.read_meta(db + ".meta")
        assert (meta.kmerlen, meta.sketchlen, meta.winlen, meta.winstride) == (k, s, w, st)
        part = dbformat.read_cache(db + ".cache0")
        out[f"g{gi}_keys"], out[f"g{gi}_sizes"], out[f"g{gi}_values"] = part.keys, part.sizes, part.values
        ob = os.path.join(tmp, f"o{gi}.bin")
        refio.run_harness(db, rt, ob, maxcand=2)
        out.update(harness_to_arrays(refio.parse_harness_output(ob), f"g{gi}_"))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "g3.npz"), **out)
    print("g3.npz", os.path.getsize(os.path.join(ROOT, "tests", "golden", "g3.npz")), "bytes,", len(pick), "reads x", len(GEOMETRIES), "geometries")


if __name__ == "__main__":
    main()
