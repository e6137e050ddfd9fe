"""CPU-side helpers of bench.py (the measured paths themselves need a GPU)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_write_reads_txt_variable_lengths(tmp_path):
    reads = [b"ACGT", b"A", b"", b"GGGTTTNN"]
    flat = np.frombuffer(b"".join(reads), np.uint8)
    offs = np.cumsum([0] + [len(r) for r in reads])
    p = tmp_path / "r.txt"
    bench.write_reads_txt(str(p), flat, offs)
    assert open(p, "rb").read() == b"ACGT\nA\n\nGGGTTTNN\n"
    # offsets that do not start at zero (a slice of a larger batch)
    bench.write_reads_txt(str(p), flat[4:], offs[1:])
    assert open(p, "rb").read() == b"A\n\nGGGTTTNN\n"


def test_bench_refuses_to_run_without_a_gpu():
    """no CPU fallback: the benchmark of the product path must not silently run on the host"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        return
    assert out.returncode != 0 and "CUDA" in (out.stderr + out.stdout)


def test_committed_bench_lines_follow_the_contract():
    """profiles/bench_*.json (the lines the round's numbers come from) carry every key of the contract"""
    prof = os.path.join(ROOT, "profiles")
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"}
    seen = 0
    for f in sorted(os.listdir(prof)):
        if not (f.startswith("bench_") and f.endswith(".json")) or f.startswith("bench_ref"):
            continue
        d = json.load(open(os.path.join(prof, f)))
        assert need <= set(d), (f, need - set(d))
        assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
        assert "workload" in d["config"] and "model" not in d["config"]
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        seen += 1
    assert seen >= 3
    for f in sorted(os.listdir(prof)):
        if f.startswith("bench_ref") and f.endswith(".json"):
            d = json.load(open(os.path.join(prof, f)))
            assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference"
            assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_shard_mode_follows_the_memory_of_the_gpu():
    """--shard-by auto: all parts merged on every GPU while the N-part database fits one B200, else the
    target-sharded capacity mode (BASELINE config C5 as specified: 8 x 1.5 G locations)"""
    import argparse
    b200 = 178 * 2 ** 30

    def args(**kw):
        d = dict(shard_by="auto", replicate_merged=False, merged_parts=0, replicate=False, targets=50_000, target_len=100_000)
        d.update(kw)
        return argparse.Namespace(**d)
    assert bench.choose_shard_mode(args(), 1, b200) == "single"
    assert [bench.choose_shard_mode(args(), n, b200) for n in (2, 4, 8)] == ["merged"] * 3
    assert bench.choose_shard_mode(args(targets=105_000), 8, b200) == "target"          # C5 does not fit merged
    assert bench.choose_shard_mode(args(), 8, 80 * 2 ** 30) == "target"                 # a smaller GPU
    assert bench.choose_shard_mode(args(shard_by="target"), 4, b200) == "target"
    assert bench.choose_shard_mode(args(shard_by="feature"), 4, b200) == "feature"
    assert bench.choose_shard_mode(args(replicate=True), 4, b200) == "single"
    assert bench.choose_shard_mode(args(merged_parts=8), 1, b200) == "merged"
