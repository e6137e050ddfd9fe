"""Generates tests/golden/*.npz from the REFERENCE ITSELF - TEST INFRASTRUCTURE ONLY.

Runs only in the build container (needs /root/reference and oracle/_ref built by
oracle/Makefile).  Everything it writes into tests/golden/ is committed so that
the GPU box (which has no /root/reference) can check parity against reference
outputs:

  g1.npz   small single-part database built by the reference `metacache build`
           (16 targets: slices of the bundled genomes, mutated copies -> ties and
           multi-target hits, and a tandem-repeat target -> buckets capped at 254)
           + 700 single / 150 paired reads incl. edge cases, with the reference's
           sketches, sorted all-hits and top candidates (maxcand 2 and 5,
           insert sizes 0 and 1000) from oracle/ref_harness.cpp
  g2.npz   the same targets built with `-parts 2`; per-part reference outputs
           (`part=0`, `part=1`); the all-parts reference run is NOT used
  kat.npz  known-answer vectors of SURVEY.md 4.3 (generated from the reference headers)

Also fills oracle/_ref/c1/ (git-ignored, travels with gpurun) with the bundled
test database `bacteria1` built by the reference, the bundled reads and
`classified.expected`, for the full C1 parity test.

usage: python oracle/make_golden.py
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tarfile
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refio  # noqa: E402
from metacache_b200 import dbformat  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
C1 = os.path.join(ROOT, "oracle", "_ref", "c1")


def mutate(rng, seq: bytes, rate: float, n_rate: float = 0.0) -> bytes:
    a = np.frombuffer(seq, dtype=np.uint8).copy()
    m = rng.random(len(a)) < rate
    a[m] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(m.sum()))
    if n_rate:
        a[rng.random(len(a)) < n_rate] = ord("N")
    return a.tobytes()


def revcomp(s: bytes) -> bytes:
    return s.translate(bytes.maketrans(b"ACGTacgt", b"TGCAtgca"))[::-1]


def pack_ragged(arrs, dtype):
    off = np.zeros(len(arrs) + 1, dtype=np.int64)
    for i, a in enumerate(arrs):
        off[i + 1] = off[i] + len(a)
    flat = np.concatenate([np.asarray(a, dtype=dtype) for a in arrs]) if len(arrs) and off[-1] else np.zeros(0, dtype)
    return flat, off


def harness_to_arrays(res, prefix):
    """list of per-read dicts -> flat arrays for npz"""
    out = {}
    sk_flat, sk_off, sk_cnt = [], [0], []
    for r in res:
        sk_cnt.append(len(r["sketches"]))
        for s in r["sketches"]:
            sk_flat.append(s)
            sk_off.append(sk_off[-1] + len(s))
    out[prefix + "sk_feats"] = np.concatenate(sk_flat).astype(np.uint32) if sk_flat and sk_off[-1] else np.zeros(0, np.uint32)
    out[prefix + "sk_off"] = np.asarray(sk_off, np.int64)
    out[prefix + "sk_cnt"] = np.asarray(sk_cnt, np.int32)
    ah, aho = pack_ragged([r["allhits"] for r in res], np.uint64)
    out[prefix + "allhits"], out[prefix + "allhits_off"] = ah, aho
    tp, tpo = pack_ragged([np.asarray(r["top"], np.uint32).reshape(-1, 4) for r in res], np.uint32)
    out[prefix + "top"] = tp.reshape(-1, 4)
    out[prefix + "top_off"] = tpo
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "all"], stdout=subprocess.DEVNULL)
    tmp = tempfile.mkdtemp(prefix="mcgold_")
    with tarfile.open(os.path.join(REF, "test", "data.tar.gz")) as t:
        t.extractall(tmp)
    with tarfile.open(os.path.join(REF, "test", "taxonomy.tar.gz")) as t:
        t.extractall(tmp)
    rng = np.random.default_rng(20260117)
    genomes = refio.read_fasta(os.path.join(tmp, "data", "bacteria1.fa"))

    # ---------------- g1 targets ----------------
    targets = []
    for i, (h, s) in enumerate(genomes[:10]):
        o = 1000 + 137 * i
        targets.append((f"G{i}", s[o:o + 12000].upper()))
    for i in (0, 3, 7):
        base = targets[i][1]
        targets.append((f"G{i}m1", mutate(rng, base, 0.01)))
        targets.append((f"G{i}m2", mutate(rng, base, 0.03)))
    unit = genomes[11][1][5000:5112].upper()
    assert len(unit) == 112
    targets.append(("REP", unit * 300))       # every window identical -> buckets hit the 254 cap
    fa = os.path.join(tmp, "g1.fa")
    with open(fa, "wb") as f:
        for name, s in targets:
            f.write(b">" + name.encode() + b"\n")
            for j in range(0, len(s), 80):
                f.write(s[j:j + 80] + b"\n")
    mc = refio.METACACHE
    subprocess.check_call([mc, "build", os.path.join(tmp, "g1"), fa, "-parts", "1", "-silent"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    # parts are only created per input FILE (building.cpp:611-616): one file per target
    os.makedirs(os.path.join(tmp, "g2in"))
    for i, (name, s) in enumerate(targets):
        with open(os.path.join(tmp, "g2in", f"t{i:02d}.fa"), "wb") as f:
            f.write(b">" + name.encode() + b"\n" + s + b"\n")
    subprocess.check_call([mc, "build", os.path.join(tmp, "g2"), os.path.join(tmp, "g2in"), "-parts", "2",
                           "-threads", "2", "-silent"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)

    # ---------------- reads ----------------
    singles = []
    tseqs = [s for _, s in targets]
    for _ in range(420):
        t = tseqs[rng.integers(len(tseqs))]
        ln = int(rng.choice([50, 75, 100, 126, 127, 128, 150, 150, 150, 200, 239, 240, 251, 400]))
        o = int(rng.integers(0, len(t) - ln))
        r = mutate(rng, t[o:o + ln], 0.01, 0.001)
        if rng.random() < 0.5:
            r = revcomp(r)
        singles.append(r)
    for _ in range(40):                                    # long reads
        t = tseqs[rng.integers(len(tseqs) - 1)]
        ln = int(rng.integers(600, 5000))
        o = int(rng.integers(0, len(t) - ln))
        singles.append(mutate(rng, t[o:o + ln], 0.05))
    singles.append(mutate(rng, tseqs[-1][100:19100], 0.002))   # 19 kbp read in the repeat target
    for _ in range(60):                                    # random reads (mostly no hits)
        singles.append(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=150).tobytes())
    bundled = refio.read_fasta(os.path.join(tmp, "data", "single.fa"))
    singles += [s for _, s in bundled[:150]]
    edge = [b"", b"A", b"ACGTACGTACGTACG", b"ACGTACGTACGTACGT", b"ACGTACGTACGTACGTA", b"N" * 150,
            b"ACGT" * 40, b"A" * 150, tseqs[0][:127], tseqs[0][:128], tseqs[0][:239], tseqs[0][:240],
            tseqs[0][:112 + 15], tseqs[0][:112 + 16], tseqs[0][100:250].lower(),
            tseqs[0][100:250].replace(b"T", b"U"), tseqs[0][100:180] + b"NNNN" + tseqs[0][184:260],
            tseqs[0][300:450] + b"RYKM", tseqs[1][:150].replace(b"A", b"a"), b"ACGTNNNNACGTACGTACGTACGTAC",
            b"acgtuacgtuacgtuacgtu", tseqs[-1][:150], tseqs[-1][:300], tseqs[-1][56:206]]
    singles += edge
    pairs = []
    for _ in range(130):
        t = tseqs[rng.integers(len(tseqs))]
        o = int(rng.integers(0, len(t) - 500))
        l1, l2 = int(rng.choice([75, 100, 150])), int(rng.choice([75, 100, 150]))
        a = mutate(rng, t[o:o + l1], 0.01)
        b = revcomp(mutate(rng, t[o + 300:o + 300 + l2], 0.01))
        pairs.append((a, b))
    pairs += [(b"", tseqs[2][:150]), (tseqs[2][:150], b""), (b"ACGT", b"ACG"), (tseqs[4][:100], b"N" * 50),
              (tseqs[0][:150], tseqs[5][:150])]
    bp = refio.read_fasta(os.path.join(tmp, "data", "pairs.fa"))
    pairs += [(bp[i][1], bp[i + 1][1]) for i in range(0, 30, 2)]
    reads = list(singles) + list(pairs)
    # harness input: whitespace separated, "-" = empty sequence
    def norm(x):
        return x if len(x) else b"-"
    reads_txt = [norm(r) if isinstance(r, bytes) else (norm(r[0]), norm(r[1])) for r in reads]
    rt = os.path.join(tmp, "reads.txt")
    refio.write_reads_txt(rt, reads_txt)

    out = {}
    meta = dbformat.read_meta(os.path.join(tmp, "g1.meta"))
    part = dbformat.read_cache(os.path.join(tmp, "g1.cache0"))
    out["keys"], out["sizes"], out["values"] = part.keys, part.sizes, part.values
    out["sketching"] = np.asarray([meta.kmerlen, meta.sketchlen, meta.winlen, meta.winstride], np.uint32)
    out["target_windows"] = meta.target_windows()
    out["target_names"] = np.asarray(meta.target_names())
    tflat, toff = pack_ragged([np.frombuffer(s, np.uint8) for _, s in targets], np.uint8)
    out["targets"], out["targets_off"] = tflat, toff
    flat1, off1 = pack_ragged([np.frombuffer(r if isinstance(r, bytes) else r[0], np.uint8) for r in reads], np.uint8)
    flat2, off2 = pack_ragged([np.frombuffer(b"" if isinstance(r, bytes) else r[1], np.uint8) for r in reads], np.uint8)
    out["reads1"], out["reads1_off"], out["reads2"], out["reads2_off"] = flat1, off1, flat2, off2
    for tag, kw in (("c2_", dict(maxcand=2)), ("c5_", dict(maxcand=5)), ("c2i_", dict(maxcand=2, insert=1000))):
        ob = os.path.join(tmp, tag + "out.bin")
        refio.run_harness(os.path.join(tmp, "g1"), rt, ob, **kw)
        out.update(harness_to_arrays(refio.parse_harness_output(ob), tag))
    np.savez_compressed(os.path.join(GOLD, "g1.npz"), **out)

    # ---------------- g2: two parts ----------------
    out2 = {}
    meta2 = dbformat.read_meta(os.path.join(tmp, "g2.meta"))
    assert meta2.num_parts == 2
    out2["target_names"] = np.asarray(meta2.target_names())
    for p in (0, 1):
        cp = dbformat.read_cache(os.path.join(tmp, f"g2.cache{p}"))
        out2[f"p{p}_keys"], out2[f"p{p}_sizes"], out2[f"p{p}_values"] = cp.keys, cp.sizes, cp.values
        ob = os.path.join(tmp, f"g2p{p}.bin")
        refio.run_harness(os.path.join(tmp, "g2"), rt, ob, part=p, maxcand=2)
        out2.update(harness_to_arrays(refio.parse_harness_output(ob), f"p{p}_"))
    np.savez_compressed(os.path.join(GOLD, "g2.npz"), **out2)

    # ---------------- KATs (values printed by the reference headers) ----------------
    np.savez_compressed(
        os.path.join(GOLD, "kat.npz"),
        hash_in=np.asarray([0, 1, 0xFFFFFFFF, 0x12345678], np.uint32),
        hash_out=np.asarray([0, 824515495, 539527247, 89967310], np.uint32),
        S=np.frombuffer(b"ACGTACGTTGCAAGCTTAGCCGATCGATTAGCNACGATCGGCTAGCTAGGATCGATCGTAGCTAGCTAGCATCGATCGATGCTAGCTAGCTAGCATGCATGCATCGATGCATGCATGCTAGTCGATGCATGCTAGTCAGTCGATGCTAGCTGATCGTAGCTAGCTAGCTGACTGATCGTAGCTAGCTAGTCGATCG", np.uint8),
        S_win0=np.asarray([8261467, 31870809, 42787827, 111240054, 163288426, 188766765, 268173904, 306425076, 380750536, 472954415, 476186958, 587694004, 601906040, 607726054, 645425469, 660504352], np.uint32),
        S_win1=np.asarray([76395789, 97501074, 238047221, 322114395, 449826254, 490607833, 537794333, 619003677, 643902065, 656383844, 656715527, 782176155, 786924561, 835612228, 854618109, 916709753], np.uint32),
        amb_seq=np.frombuffer(b"ACGTNNNNACGTACGTACGTACGTAC", np.uint8),
        amb_feats=np.asarray([1038204269, 1989685073, 3480792923], np.uint32),
        low_seq=np.frombuffer(b"acgtuacgtuacgtuacgtu", np.uint8),
        low_feats=np.asarray([262155257, 783024520, 2244369827, 2621673391, 3154704415], np.uint32),
    )

    # ---------------- C1: bundled database + reads + reference golden file ----------------
    os.makedirs(C1, exist_ok=True)
    cwd = os.getcwd()
    os.chdir(tmp)        # the reference stores the relative input path in .meta
    subprocess.check_call([mc, "build", "bacteria1", "data/bacteria1.fa", "-taxonomy", "taxonomy", "-parts", "1", "-silent"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    os.chdir(cwd)
    for f in ("bacteria1.meta", "bacteria1.cache0"):
        shutil.copy(os.path.join(tmp, f), os.path.join(C1, f))
    for f in ("single.fa", "pairs.fa", "pair.1.fa", "pair.2.fa", "classified.expected"):
        shutil.copy(os.path.join(tmp, "data", f), os.path.join(C1, f))
    # what the UNMODIFIED CPU reference built from these sources prints for its own test (test/run_tests:140-168).
    # Upstream's classified.expected predates two formatting changes of printing.cpp (15-digit fractional
    # abundances, an extra "--" column in the "unclassified" abundance row: 20 of 36 009 lines), so the drop-in
    # CLI test compares with this capture exactly and with the upstream file everywhere else.
    write_cli_capture(mc, tmp)
    shutil.rmtree(tmp)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


CLI_COMMON = ("-no-query-params -mapped-only -precision -ground-truth -tophits -allhits -hits-per-ref "
              "-abundances -abundance-per species -threads 8")


def write_cli_capture(mc, tmp):
    """oracle/_ref/c1/cli_cpu_reference.out: stdout of the CPU reference `metacache query` for the three
    FASTA inputs of the reference's own test, run in `tmp` (which holds data/ and the bacteria1 database)"""
    q = (f"data/single.fa {CLI_COMMON}\ndata/pairs.fa -pairseq {CLI_COMMON}\n"
         f"data/pair.1.fa data/pair.2.fa -pairfiles {CLI_COMMON}\n")
    out = subprocess.run([mc, "query", "bacteria1"], input=q, capture_output=True, text=True, cwd=tmp, check=True).stdout
    with open(os.path.join(C1, "cli_cpu_reference.out"), "w") as f:
        f.write(out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cli-capture":
        # only the capture, into an existing oracle/_ref/c1 (needs the reference's test data unpacked)
        import tarfile
        t = tempfile.mkdtemp()
        tarfile.open(os.path.join(REF, "test", "data.tar.gz")).extractall(t)
        for f in ("bacteria1.meta", "bacteria1.cache0"):
            shutil.copy(os.path.join(C1, f), os.path.join(t, f))
        write_cli_capture(refio.METACACHE, t)
        shutil.rmtree(t)
    else:
        main()
