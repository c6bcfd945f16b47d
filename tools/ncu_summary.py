"""Developer tool: condense an `ncu --set full` report (.ncu-rep) into the JSON kept under profiles/ (DRAM bytes, pipe
utilisation, stall cycles per issued instruction, instruction counts) -- one entry per captured launch.

    python tools/ncu_summary.py gpurun_out/r2z_scan.ncu-rep profiles/r2z_ncu_summary.json [key=value ...]
"""
import csv
import json
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    extra = dict(kv.split("=", 1) for kv in sys.argv[3:])
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        e = {"kernel": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"]}
        for k in hdr:
            if k in KEEP or "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                if "issue_stalled" in k and v < 0.05:
                    continue
                e[k] = f"{d[k]} {u[k]}".strip()
        launches.append(e)
    json.dump({"source": rep, **extra, "launches": launches}, open(out, "w"), indent=1)
    print(out, len(launches), "launches")


if __name__ == "__main__":
    main()
