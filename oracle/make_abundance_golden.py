"""tests/golden/abundance.json: the abundance tables the UNMODIFIED CPU reference prints
(`metacache query ... -abundances -abundance-per <rank>`, classification.cpp:304-377, printing.cpp:424-497)
for its own test inputs at several estimation ranks.  Run in the container that has /root/reference, after
__graft_entry__.build() (needs oracle/_ref/metacache and oracle/_ref/c1)."""
import json
import os
import shutil
import subprocess
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refio                                   # noqa: E402
from tests.golden_util import C1, reference_abundance_blocks   # noqa: E402

RANKS = ["sequence", "subspecies", "species", "genus", "family", "order", "phylum", "domain"]
INPUTS = {"single": "data/single.fa", "pairs": "data/pairs.fa -pairseq"}


def main():
    tmp = tempfile.mkdtemp()
    tarfile.open(os.path.join("/root/reference", "test", "data.tar.gz")).extractall(tmp)
    for f in ("bacteria1.meta", "bacteria1.cache0"):
        shutil.copy(os.path.join(C1, f), os.path.join(tmp, f))
    out = {}
    for rank in RANKS:
        q = "".join(f"{inp} -no-query-params -mapped-only -abundances -abundance-per {rank} -threads 4\n" for inp in INPUTS.values())
        txt = subprocess.run([refio.METACACHE, "query", "bacteria1"], input=q, capture_output=True, text=True, cwd=tmp, check=True).stdout
        cap = os.path.join(tmp, "cap.out")
        open(cap, "w").write(txt)
        blocks = reference_abundance_blocks(cap)
        assert len(blocks) == len(INPUTS), (rank, len(blocks))
        out[rank] = dict(zip(INPUTS, blocks))
    shutil.rmtree(tmp)
    dst = os.path.join(ROOT, "tests", "golden", "abundance.json")
    json.dump({"generated_by": "oracle/make_abundance_golden.py", "reference": "muellan/metacache @ d7646ec, CPU build", "blocks": out},
              open(dst, "w"), indent=1)
    print(dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
