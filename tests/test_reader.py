"""FASTA/FASTQ reader (csrc/reader.cpp) against the reference's own reader: the golden
records in tests/golden/reader.json were produced by oracle/_ref/mc_ref_reader (ref_reader.cpp linked
with the unmodified sequence_io.cpp) from the inputs in tests/reader_cases.py."""
import gzip
import json
import os
import zlib

import numpy as np
import pytest

from metacache_b200._lib import Mcb200Error
from metacache_b200.reader import SequenceReader
from tests.reader_cases import CASES, big_content

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = {e["name"]: e for e in json.load(open(os.path.join(HERE, "golden", "reader.json")))}


def _write(case, d):
    paths = []
    for i, (kind, data) in enumerate(case["files"]):
        raw = big_content(data) if kind == "big" else data
        if case.get("gz"):
            raw = gzip.compress(raw)
        p = os.path.join(d, f"{case['name']}_{i}" + (".gz" if case.get("gz") else ""))
        open(p, "wb").write(raw)
        paths.append(p)
    if case.get("pairseq"):
        paths = [paths[0], paths[0]]
    return paths


def _records(reader):
    out = []
    for h, a, b in reader:
        out.append([reader.index(), h.decode("latin1"), a.decode("latin1"), b.decode("latin1")])
    return out


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reader_matches_reference_reader(case, tmp_path):
    gold = GOLD[case["name"]]
    paths = _write(case, str(tmp_path))
    if "error" in gold:
        with pytest.raises(Mcb200Error) as ei:
            SequenceReader(*paths)
        assert gold["error"] in str(ei.value)
        return
    with SequenceReader(*paths) as r:
        recs = _records(r)
    if "digest" in gold:
        g = gold["digest"]
        assert len(recs) == g["n"]
        assert [len(x[2]) for x in recs[:50]] == g["lens"]
        assert zlib.crc32(json.dumps(recs, separators=(",", ":")).encode()) & 0xFFFFFFFF == g["crc"]
    else:
        assert recs == gold["records"]


@pytest.mark.parametrize("name", ["big_fasta_long_lines", "big_fasta_wrapped", "big_fastq", "fasta_simple",
                                  "fastq_qual_starts_with_at", "fasta_no_final_newline"])
@pytest.mark.parametrize("nranges", [2, 3, 7, 64])
def test_byte_ranges_partition_the_records(name, nranges, tmp_path):
    """readers over disjoint byte ranges deliver every record exactly once, in file order"""
    case = next(c for c in CASES if c["name"] == name)
    path = _write(case, str(tmp_path))[0]
    with SequenceReader(path) as r:
        whole = [(h, a) for h, a, _ in r]
    size = os.path.getsize(path)
    rng = np.random.default_rng(nranges)
    cuts = sorted(set([0, size] + [int(x) for x in rng.integers(0, size + 1, nranges - 1)]))
    got = []
    for b, e in zip(cuts[:-1], cuts[1:]):
        with SequenceReader(path, byte_range=(b, e)) as r:
            got += [(h, a) for h, a, _ in r]
    assert got == whole


def test_bundled_reference_reads_if_present():
    """the reads of the reference's own test suite (present where oracle/_ref/c1 was unpacked)"""
    from tests.golden_util import need_c1
    need_c1()
    c1 = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "c1")
    with SequenceReader(os.path.join(c1, "single.fa")) as r:
        n = sum(1 for _ in r)
    heads = sum(1 for line in open(os.path.join(c1, "single.fa"), "rb") if line.startswith(b">"))
    assert n == heads and n > 1000
    with SequenceReader(os.path.join(c1, "pair.1.fa"), os.path.join(c1, "pair.2.fa")) as r:
        pairs = list(r)
    assert all(len(a) and len(b) for _, a, b in pairs)


def test_reader_errors():
    with pytest.raises(Mcb200Error):
        SequenceReader("/nonexistent/file.fa")


def _dump_exe():
    import subprocess
    host = os.path.join(os.path.dirname(HERE), "metacache_b200", "host")
    exe = os.path.join(host, "shim_reader_dump")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", host, "shim_reader_dump"], stdout=subprocess.DEVNULL)
    return exe


def _parse_dump(out):
    recs = []
    for line in out.split(b"\n"):
        if not line:
            continue
        f = line.split(b"\t")
        if f[0] == b"ERROR":
            return {"error": f[1].decode()}
        recs.append([int(f[0]), f[1].decode("latin1"), f[2].decode("latin1"), f[3].decode("latin1")])
    return {"records": recs}


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cpp_reader_mirror_matches_reference_reader(case, tmp_path):
    """mcb200::sequence_pair_reader (host/mcb200_shim.hpp: has_next / next / index over the C ABI) prints
    the same lines as the reference's sequence_pair_reader (oracle/ref_reader.cpp)"""
    import subprocess
    gold = GOLD[case["name"]]
    paths = _write(case, str(tmp_path))
    res = _parse_dump(subprocess.run([_dump_exe()] + paths, capture_output=True).stdout)
    if "error" in gold:
        assert gold["error"] in res.get("error", "")
        return
    recs = res["records"]
    if "digest" in gold:
        g = gold["digest"]
        assert len(recs) == g["n"] and [len(x[2]) for x in recs[:50]] == g["lens"]
        assert zlib.crc32(json.dumps(recs, separators=(",", ":")).encode()) & 0xFFFFFFFF == g["crc"]
    else:
        assert recs == gold["records"]


def test_cpp_reader_skip(tmp_path):
    """sequence_pair_reader::skip(n): the remaining queries and their indices are unchanged"""
    import subprocess
    case = next(c for c in CASES if c["name"] == "big_fastq")
    path = _write(case, str(tmp_path))[0]
    whole = _parse_dump(subprocess.run([_dump_exe(), path], capture_output=True).stdout)["records"]
    for k in (0, 1, 7, 19999, 20000, 30000):
        part = _parse_dump(subprocess.run([_dump_exe(), path, "", str(k)], capture_output=True).stdout)["records"]
        assert part == whole[k:]


def _multiline_fastq(n, seed):
    """FASTQ with multi-line sequences and ONE quality line per record (the grammar the reference
    reader accepts, sequence_io.cpp:203-223); quality lines start with '@' every few records."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        ln = int(rng.integers(0, 400))
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), ln))
        width = int(rng.integers(20, 90))
        lines = [seq[j:j + width] for j in range(0, ln, width)] or [b""]
        if i % 11 == 3:
            lines = lines[:1] + [b""] + lines[1:]                      # an empty line inside the sequence
        qual = bytes(rng.integers(33, 74, ln, dtype=np.uint8))
        if i % 3 == 0 and ln:
            qual = b"@" + qual[1:]
        out.append(b"@r%d some text\n" % i + b"\n".join(lines) + b"\n+" + (b"r%d" % i if i % 2 else b"") + b"\n" + qual + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("nranges", [2, 4, 7, 33])
def test_byte_ranges_on_multiline_fastq(nranges, tmp_path):
    """range readers must resynchronise on FASTQ records with multi-line sequences too"""
    path = os.path.join(str(tmp_path), "ml.fq")
    open(path, "wb").write(_multiline_fastq(400, 5))
    with SequenceReader(path) as r:
        whole = [(h, a) for h, a, _ in r]
    assert len(whole) == 400
    size = os.path.getsize(path)
    rng = np.random.default_rng(100 + nranges)
    cuts = sorted(set([0, size] + [int(x) for x in rng.integers(0, size + 1, nranges - 1)]))
    got = []
    for b, e in zip(cuts[:-1], cuts[1:]):
        with SequenceReader(path, byte_range=(b, e)) as r:
            got += [(h, a) for h, a, _ in r]
    assert got == whole
    # every possible cut position of a small file
    small = os.path.join(str(tmp_path), "ml_small.fq")
    open(small, "wb").write(_multiline_fastq(6, 9))
    with SequenceReader(small) as r:
        whole = [(h, a) for h, a, _ in r]
    size = os.path.getsize(small)
    for cut in range(0, size + 1):
        got = []
        for b, e in ((0, cut), (cut, size)):
            with SequenceReader(small, byte_range=(b, e)) as r:
                got += [(h, a) for h, a, _ in r]
        assert got == whole, cut


def test_truncated_gzip_is_an_error_not_an_end_of_input(tmp_path):
    """a corrupt / truncated gzip stream must not look like a clean end of the reads"""
    raw = b"".join(b">r%d\n%s\n" % (i, b"ACGT" * 40) for i in range(20000))
    z = gzip.compress(raw)
    path = os.path.join(str(tmp_path), "trunc.fa.gz")
    open(path, "wb").write(z[:len(z) // 2])
    with pytest.raises(Mcb200Error):
        with SequenceReader(path) as r:
            n = sum(1 for _ in r)
            raise AssertionError(f"read {n} records from a truncated gzip file without an error")
