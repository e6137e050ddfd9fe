"""tests/golden/classify_options.json: per-read `top_hits` and classification columns the UNMODIFIED CPU
reference prints for its own test reads under NON-default candidate / classification options
(-hitmin, -hitdiff, -maxcand, -lowest, -highest: options.cpp:865-915; classify(), classification.cpp:146-189;
insert() with taxon merge, candidate_generation.hpp:172-231; show_candidates, printing.cpp:283-310).
Run in the container that has /root/reference, after __graft_entry__.build()."""
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
from tests.golden_util import C1                           # noqa: E402

# name -> (CLI options, the same as oracle / device parameters)
SETS = {
    "hitmin3_diff0.5_cand4": ("-hitmin 3 -hitdiff 0.5 -maxcand 4", dict(hits_min=3, frac=0.5, maxc=4, lowest=0, highest=19)),
    "hitmin8_diff0.2_family_cand3": ("-hitmin 8 -hitdiff 0.2 -highest family -maxcand 3", dict(hits_min=8, frac=0.2, maxc=3, lowest=0, highest=10)),
    "lowest_species_cand3": ("-lowest species -maxcand 3", dict(hits_min=5, frac=1.0, maxc=3, lowest=4, highest=19)),
    "lowest_genus_order_hitmin2_cand5": ("-lowest genus -highest order -hitmin 2 -hitdiff 1 -maxcand 5", dict(hits_min=2, frac=1.0, maxc=5, lowest=6, highest=12)),
    "diff0_cand2": ("-hitdiff 0 -maxcand 2", dict(hits_min=5, frac=0.0, maxc=2, lowest=0, highest=19)),
}
INPUTS = {"single": "data/single.fa", "pairs": "data/pairs.fa -pairseq"}


def main():
    tmp = tempfile.mkdtemp()
    tarfile.open(os.path.join("/root/reference", "test", "data.tar.gz")).extractall(tmp)
    for f in ("bacteria1.meta", "bacteria1.cache0"):
        shutil.copy(os.path.join(C1, f), os.path.join(tmp, f))
    out = {}
    for name, (cli, params) in SETS.items():
        out[name] = {"cli": cli, "params": params, "reads": {}}
        for inp, path in INPUTS.items():
            q = f"{path} -no-query-params -no-summary -mapped-only -tophits -queryids -threads 1 {cli}\n"
            txt = subprocess.run([refio.METACACHE, "query", "bacteria1"], input=q, capture_output=True, text=True,
                                 cwd=tmp, check=True).stdout
            rows = {}
            for line in txt.splitlines():
                cols = line.split("\t|\t")
                if len(cols) == 4 and cols[0].isdigit():
                    rows[cols[0]] = [cols[2], cols[3]]              # query id -> top_hits, classification
            assert rows or inp != "single", (name, inp)
            out[name]["reads"][inp] = rows
    shutil.rmtree(tmp)
    dst = os.path.join(ROOT, "tests", "golden", "classify_options.json")
    json.dump({"generated_by": "oracle/make_classify_golden.py", "reference": "muellan/metacache @ d7646ec, CPU build",
               "note": "-mapped-only: reads that are missing were not classified", "sets": out}, open(dst, "w"), indent=0)
    print(dst, os.path.getsize(dst), "bytes", {k: {i: len(v) for i, v in s["reads"].items()} for k, s in out.items()})


if __name__ == "__main__":
    main()
