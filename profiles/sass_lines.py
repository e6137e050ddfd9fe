"""Joins an `ncu --page source --csv` SASS export with nvdisasm line info:
per source line, instructions executed per unit (read/window) and stall samples.

usage: sass_lines.py <src.csv> <nvdisasm -g -c output> <mangled function> <units> [top]
"""
import collections
import csv
import re
import sys


def main(src_csv, sass_txt, func, units, top=40):
    lines = open(sass_txt).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + func + ":"))
    cur, seq = None, []
    for l in lines[start + 1:]:
        if l.startswith("//--------------------- .text."):
            break
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            seq.append(cur)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    print(f"# {len(seq)} SASS instructions with line info, {len(data)} in the profile")
    agg, samp = collections.Counter(), collections.Counter()
    for loc, r in zip(seq, data):
        agg[loc] += int(r[idx["Instructions Executed"]]) / units
        samp[loc] += int(r[idx["# Samples"]])
    tot_s = sum(samp.values())
    print(f"# total {sum(agg.values()):.1f} warp-instructions per unit, {tot_s} samples")
    cache = {}
    for loc, n in sorted(agg.items(), key=lambda x: -samp[x[0]])[:top]:
        f, ln = loc if loc else ("?", 0)
        text = ""
        if ln:
            for d in ("/root/repo/metacache_b200/csrc/",):
                try:
                    cache.setdefault(f, open(d + f).read().splitlines())
                    text = cache[f][ln - 1].strip()[:80]
                except OSError:
                    pass
        print(f"{n:8.1f} inst/unit {100.0 * samp[loc] / tot_s:5.1f}% smp  {f}:{ln}  {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), int(sys.argv[5]) if len(sys.argv) > 5 else 40)
