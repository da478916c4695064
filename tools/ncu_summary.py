"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small text summary
for profiles/.   usage: python tools/ncu_summary.py report.ncu-rep [out.txt]"""

import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "gpc__cycles_elapsed.avg.per_second",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__cycles_active.avg",
]


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"== kernel: {d.get('Kernel Name')}  (ID {d.get('ID')})  grid {d.get('Grid Size')} block {d.get('Block Size')}", file=out)
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>18s} {u.get(k, '')}", file=out)
        stalls = [(k, float(d[k])) for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k]]
        stalls.sort(key=lambda kv: -kv[1])
        print("  warp stall reasons (warps stalled per issue-active cycle), top 8:", file=out)
        for k, v in stalls[:8]:
            name = k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
            print(f"    {name:40s} {v:8.3f}", file=out)


if __name__ == "__main__":
    main()
