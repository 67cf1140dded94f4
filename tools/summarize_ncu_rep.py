"""Summarise `ncu --set full` reports (.ncu-rep) into the text format bench.py parses (profiles/*.summary.txt):
    python tools/summarize_ncu_rep.py OUT.txt  rep1.ncu-rep "label 1"  rep2.ncu-rep "label 2" ..."""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__cluster_dim_x",
        "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    out, rest = sys.argv[1], sys.argv[2:]
    lines = []
    for rep, label in zip(rest[0::2], rest[1::2]):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        name = vals[hdr.index("Kernel Name")][:70]
        lines.append(f"== {name}  {label}   [ncu --set full --clock-control none, 1 launch]")
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"   {k:70s} {vals[i]} {units[i]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
