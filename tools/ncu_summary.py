"""Condense an .ncu-rep (ncu --set full) into the small JSON summaries kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_ncu_summary.json "command that produced it"
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
]


def main():
    rep, out, command = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(header)}
    kernels = []
    for r in rows[2:]:
        ent = {"kernel": r[col["Kernel Name"]], "metrics": {}}
        for m in KEEP:
            if m in col:
                ent["metrics"][m] = {"unit": units[col[m]], "value": r[col[m]]}
        kernels.append(ent)
    json.dump({"report": rep.split("/")[-1], "command": command, "kernels": kernels}, open(out, "w"), indent=1)
    print(f"{len(kernels)} kernel(s) -> {out}")


if __name__ == "__main__":
    main()
