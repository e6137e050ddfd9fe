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


def _six_digits(line):
    """The golden file predates printing.cpp:458 (`setprecision(15)` for fractional abundances): the
    unmodified CPU reference built from the same sources prints 17.7777777777778 where the file has
    17.7778 (14 lines).  Both sides are compared at the stream default of 6 significant digits."""
    return re.sub(r"\d+\.\d{7,}", lambda m: "%g" % float(m.group(0)), line)


def _filter(text):
    """run_tests:153: grep "|\\|#" | grep -v "time\\|speed\\|list\\|ignore" | sed "s/\\.fa//g" """
    out = []
    for line in text.splitlines():
        if not ("|" in line or "#" in line):
            continue
        if re.search(r"time|speed|list|ignore", line):
            continue
        out.append(_six_digits(line.replace(".fa", "")))
    return out


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
    want = sorted(_filter(open(os.path.join(C1, "classified.expected")).read()))
    assert len(want) > 36000
    if got != want:
        gs, ws = set(got), set(want)
        missing = [x for x in want if x not in gs][:5]
        extra = [x for x in got if x not in ws][:5]
        raise AssertionError(f"{len(got)} lines vs {len(want)} expected\nmissing: {missing}\nextra: {extra}\nstderr: {r.stderr[-500:]}")
