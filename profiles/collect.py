"""Turns the scratch artefacts in gpurun_out/ into the committed summaries under profiles/.

  python profiles/collect.py r1      # round tag
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def last_json(path):
    if not os.path.exists(path):
        return None
    for line in reversed(open(path, errors="replace").read().splitlines()):
        line = line.strip()
        if line.startswith("{") and line.endswith("}"):
            try:
                return json.loads(line)
            except ValueError:
                continue
    return None


def launches(tag):
    src = os.path.join(G, f"launches_{tag}.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    launch = collections.OrderedDict()
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        d = launch.setdefault(r[idx["ID"]], {"name": r[idx["Kernel Name"]].split("(")[0][:56]})
        m, v, u = r[idx["Metric Name"]], float(r[idx["Metric Value"]].replace(",", "")), r[idx["Metric Unit"]]
        if m == "gpu__time_duration.sum":
            d["ms"] = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        elif m == "dram__bytes_read.sum":
            d["rd"] = v * scale.get(u, 1)
        elif m == "dram__bytes_write.sum":
            d["wr"] = v * scale.get(u, 1)
    L = list(launch.values())
    agg = collections.OrderedDict()
    for d in L:
        a = agg.setdefault(d["name"], [0, 0.0, 0.0])
        a[0] += 1; a[1] += d.get("ms", 0); a[2] += d.get("rd", 0) + d.get("wr", 0)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"launches_{tag}_summary.txt"), "w") as f:
        f.write(f"# ncu launch list of the library's kernels during `bench.py --steps 2 --warmup 1` ({len(L)} launches;\n"
                "# cold-cache, serialised: compare SHARES).  Includes the database build and the e2e slots.\n")
        f.write(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'dram GB':>9s}\n")
        for k, (n, ms, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:58s} {n:8d} {ms:10.3f} {100 * ms / tot:6.1f}% {b / 1e9:9.3f}\n")
        f.write("\n# last query step (one 1 M-read e2e slot), in launch order\n")
        for d in L[-6:]:
            f.write(f"{d['name']:58s} {d.get('ms', 0):9.3f} ms  dram rd {d.get('rd', 0) / 1e9:7.3f} GB  wr {d.get('wr', 0) / 1e9:7.3f} GB\n")


def ncu_text(rep, out, lines_func=None, units=1e6):
    rep = os.path.join(G, rep)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    tmp = os.path.join(G, "_raw.csv")
    open(tmp, "w").write(raw)
    txt = subprocess.run([sys.executable, os.path.join(P, "ncu_summary.py"), tmp], capture_output=True, text=True).stdout
    open(os.path.join(P, out + ".txt"), "w").write(f"# from {os.path.basename(rep)} (ncu --set full --clock-control none)\n" + txt)
    if lines_func:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        tmp2 = os.path.join(G, "_src.csv")
        open(tmp2, "w").write(src)
        cub = "/tmp/cub_collect"
        os.makedirs(cub, exist_ok=True)
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "metacache_b200", "libmcb200.so")], cwd=cub,
                       capture_output=True)
        cubin = "kernels_query.sm_100a.cubin" if "query" in lines_func else "kernels_sketch.sm_100a.cubin"
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(cub, cubin)], capture_output=True, text=True).stdout
        open(os.path.join(cub, "all.sass"), "w").write(sass)
        # lines_func is a prefix of the mangled name (the parameter list changes with the kernel's signature)
        m = re.search(r"^\.text\.(" + re.escape(lines_func) + r"\w*):", sass, re.M)
        if m:
            lines_func = m.group(1)
        t = subprocess.run([sys.executable, os.path.join(P, "sass_lines.py"), tmp2, os.path.join(cub, "all.sass"),
                            lines_func, str(units), "40"], capture_output=True, text=True).stdout
        open(os.path.join(P, out.replace("ncu_", "ncu_lines_") + ".txt"), "w").write(
            f"# {os.path.basename(rep)} joined with nvdisasm -g line info; unit = one read / one window\n" + t)


def main(tag):
    for log, name in (("bench_full.log", f"bench_{tag}.json"), ("bench_ref.log", f"bench_ref_{tag}.json"),
                      ("bench_n2.log", f"bench_n2_{tag}.json"), ("bench_n4.log", f"bench_n4_{tag}.json"),
                      ("bench_n8.log", f"bench_n8_{tag}.json"), ("bench_c3.log", f"bench_c3_{tag}.json"),
                      ("bench_ref_n2.log", f"bench_ref_n2_{tag}.json"), ("bench_ref_n4.log", f"bench_ref_n4_{tag}.json"),
                      ("bench_ref_n8.log", f"bench_ref_n8_{tag}.json")):
        j = last_json(os.path.join(G, log))
        if j:
            json.dump(j, open(os.path.join(P, name), "w"), indent=1)
    launches(tag)
    ncu_text(f"prof_query_{tag}.ncu-rep", f"ncu_query_fast_{tag}", "_ZN3mcb17query_fast_kernelIjLb0ELi0E", 1e6)
    ncu_text(f"prof_sketch_{tag}.ncu-rep", f"ncu_sketch_{tag}", "_ZN3mcb18sketch_fast_kernelE", 2e6)
    for f in ("gather_bench3.log", "exp1.log"):
        if os.path.exists(os.path.join(G, f)):
            dst = os.path.join(P, "gather_bench_" + tag + ".log") if f.startswith("gather") else os.path.join(P, "exp", f"exp1_{tag}.log")
            open(dst, "w").write(open(os.path.join(G, f)).read())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1")
