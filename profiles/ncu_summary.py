"""Prints the key metrics of an `ncu --page raw --csv` export (one block per profiled launch)."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("=====", r[idx["Kernel Name"]][:100])
        for w in WANT:
            if w in idx:
                print(f"  {w:72s} {r[idx[w]]:>18s} {units[idx[w]]}")
        stalls = [(float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]) for h, i in idx.items()
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        for v, n in sorted(stalls, reverse=True)[:7]:
            print(f"  stall {n:66s} {v:18.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
