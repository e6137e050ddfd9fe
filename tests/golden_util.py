"""Loaders for tests/golden/*.npz (made by oracle/make_golden.py from the reference)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C1 = os.path.join(ROOT, "oracle", "_ref", "c1")


def _ragged(flat, off):
    return [flat[off[i]:off[i + 1]] for i in range(len(off) - 1)]


class Expected:
    """reference outputs for one harness run: per read sketches / allhits / top"""

    def __init__(self, z, prefix):
        feats, off, cnt = z[prefix + "sk_feats"], z[prefix + "sk_off"], z[prefix + "sk_cnt"]
        sk = _ragged(feats, off)
        self.sketches, p = [], 0
        for c in cnt:
            self.sketches.append(sk[p:p + c])
            p += c
        self.allhits = _ragged(z[prefix + "allhits"], z[prefix + "allhits_off"])
        top, toff = z[prefix + "top"], z[prefix + "top_off"]
        self.top = [[tuple(int(x) for x in row) for row in top[toff[i]:toff[i + 1]]] for i in range(len(toff) - 1)]


class G1:
    def __init__(self):
        z = np.load(os.path.join(GOLD, "g1.npz"))
        self.z = z
        self.keys, self.sizes, self.values = z["keys"], z["sizes"], z["values"]
        self.k, self.s, self.w, self.stride = (int(x) for x in z["sketching"])
        r1 = _ragged(z["reads1"], z["reads1_off"])
        r2 = _ragged(z["reads2"], z["reads2_off"])
        self.reads = [(a.tobytes(), b.tobytes()) for a, b in zip(r1, r2)]
        self.targets = [t.tobytes() for t in _ragged(z["targets"], z["targets_off"])]
        self.target_windows = z["target_windows"]

    def expected(self, tag):
        return Expected(self.z, tag)


class G2:
    def __init__(self):
        z = np.load(os.path.join(GOLD, "g2.npz"))
        self.z = z
        self.parts = [(z[f"p{p}_keys"], z[f"p{p}_sizes"], z[f"p{p}_values"]) for p in (0, 1)]

    def expected(self, part):
        return Expected(self.z, f"p{part}_")


def kat():
    return np.load(os.path.join(GOLD, "kat.npz"))


class G3:
    """reference outputs for non-default sketching geometries (oracle/make_golden_geometry.py)"""

    def __init__(self):
        z = np.load(os.path.join(GOLD, "g3.npz"))
        self.z = z
        g1 = G1()
        self.reads = [g1.reads[i] for i in z["read_index"]]
        self.geometries = [tuple(int(x) for x in row) for row in z["geometries"]]

    def part(self, gi):
        z = self.z
        return z[f"g{gi}_keys"], z[f"g{gi}_sizes"], z[f"g{gi}_values"]

    def expected(self, gi):
        return Expected(self.z, f"g{gi}_")


def need_c1(gpu_test=False):
    """The reference's bundled test data (oracle/_ref/c1: reference-built `bacteria1` database, reads and
    classified.expected; made by __graft_entry__.build() from /root/reference, shipped to the GPU box with
    the snapshot).  Its absence is a FAILURE wherever it can exist: in a GPU test (the artefact travels)
    and in any environment that has the reference sources; only a CPU-only checkout without the reference
    may skip."""
    import os
    import pytest
    if os.path.exists(os.path.join(C1, "classified.expected")) and os.path.exists(os.path.join(C1, "bacteria1.cache0")):
        return
    if gpu_test or os.path.exists("/root/reference/src/main.cpp"):
        pytest.fail("oracle/_ref/c1 is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where "
                    "/root/reference exists (the parity test against the reference's golden file must not vanish)")
    pytest.skip("oracle/_ref/c1 not built and no reference sources here")


def reference_abundance_blocks(path):
    """the `-abundances -abundance-per <rank>` tables in a captured output of the reference CLI: one list of
    lines per queried input (header lines, rows, the `unclassified` row; both tables)"""
    blocks, cur = [], None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith("# query summary: number of queries mapped per taxon"):
            cur = [line]
            blocks.append(cur)
        elif cur is not None:
            if line.startswith("# ") and not (line.startswith("# rank:name") or line.startswith("# estimated abundance")):
                cur = None
            elif "\t|\t" in line or line.startswith("# "):
                cur.append(line)
                if line.startswith("unclassified") and any(x.startswith("# estimated") for x in cur):
                    cur = None
    return blocks
