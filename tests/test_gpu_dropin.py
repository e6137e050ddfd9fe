"""The real drop-in: the reference's OWN `metacache` host program (CLI, options, FASTA reader, batch
executor, classification, statistics, printing) compiled unchanged with its -DGPU_MODE seam on top of
libmcb200 (oracle/Makefile: _ref/metacache_mcb200), run the way the reference's own
test runs it (test/run_tests:140-168: three FASTA inputs, single / -pairseq / -pairfiles, with
-tophits -allhits -hits-per-ref -abundances ...) and compared with the reference's CPU golden file
test/data/classified.expected - every line of the program's output, not only the candidates."""
import os
import re
import subprocess

import pytest

from tests.golden_util import C1, need_c1

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DROPIN = os.path.join(HERE, "..", "oracle", "_ref", "metacache_mcb200")
COMMON = ("-no-query-params -mapped-only -precision -ground-truth -tophits -allhits -hits-per-ref "
          "-abundances -abundance-per species -threads {threads}")


def _filter(text):
    """run_tests:153: grep "|\\|#" | grep -v "time\\|speed\\|list\\|ignore" | sed "s/\\.fa//g" """
    out = []
    for line in text.splitlines():
        if not ("|" in line or "#" in line):
            continue
        if re.search(r"time|speed|list|ignore", line):
            continue
        out.append(line.replace(".fa", ""))
    return out


def _upstream_format(line):
    """classified.expected predates two formatting changes of printing.cpp that the UNMODIFIED CPU reference
    built from the same sources shows as well (20 of 36 009 lines): fractional abundances are printed with 15
    digits (printing.cpp:458) and the "unclassified" abundance row has an extra "--" column (:462-466)."""
    line = re.sub(r"\d+\.\d{7,}", lambda m: "%g" % float(m.group(0)), line)
    return line.replace("unclassified\t|\t--\t|\t", "unclassified\t|\t")


@pytest.mark.parametrize("threads", [8, 3])
def test_reference_cli_on_libmcb200_reproduces_the_cpu_golden_file(tmp_path, threads):
    need_c1(gpu_test=True)
    assert os.path.exists(DROPIN), ("oracle/_ref/metacache_mcb200 missing: run build() where /root/reference exists "
                                    "(the binary travels to the GPU box)")
    os.symlink(C1, tmp_path / "data")
    common = COMMON.format(threads=threads)
    queries = (f"data/single.fa {common}\n"
               f"data/pairs.fa -pairseq {common}\n"
               f"data/pair.1.fa data/pair.2.fa -pairfiles {common}\n")
    r = subprocess.run([DROPIN, "query", os.path.join(C1, "bacteria1")], input=queries, capture_output=True,
                       text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = sorted(_filter(r.stdout))
    # 1. exactly what the unmodified CPU reference prints (captured by oracle/make_golden.py at build time)
    cap = os.path.join(C1, "cli_cpu_reference.out")
    assert os.path.exists(cap), "oracle/_ref/c1/cli_cpu_reference.out missing: python oracle/make_golden.py cli-capture"
    want = sorted(_filter(open(cap).read()))
    assert len(want) > 36000
    _same(got, want, r.stderr)
    # 2. the reference's own golden file, modulo the two stale formats
    want = sorted(_filter(open(os.path.join(C1, "classified.expected")).read()))
    _same(sorted(_upstream_format(x) for x in got), want, r.stderr)


def _same(got, want, stderr):
    if got != want:
        gs, ws = set(got), set(want)
        missing = [x for x in want if x not in gs][:5]
        extra = [x for x in got if x not in ws][:5]
        raise AssertionError(f"{len(got)} lines vs {len(want)} expected\nmissing: {missing}\nextra: {extra}\nstderr: {stderr[-300:]}")
