"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small JSON/text pair.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/r1_step_kernel  [particles]"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    particles = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in hdr:
            if k in KEYS or ("issue_stalled" in k and k.endswith("per_issue_active.ratio")):
                i = hdr.index(k)
                try:
                    v, u = float(r[i].replace(",", "")), units[i]
                    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u)
                    if scale is not None:
                        v, u = v * scale, "byte"
                    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u) if k.startswith("gpu__time") else None
                    if tscale is not None:
                        v, u = v * tscale, "us"
                    d[k + " [" + u + "]"] = v
                except ValueError:
                    d[k] = r[i]
        launches.append(d)
    summary = {"report": rep, "launches": launches}
    if particles and launches:
        l0 = launches[0]
        rd = l0["dram__bytes_read.sum [byte]"]
        wr = l0["dram__bytes_write.sum [byte]"]
        summary["particles"] = particles
        summary["dram_bytes_per_launch"] = rd + wr
        summary["dram_bytes_per_env_step"] = (rd + wr) / particles
        summary["warp_instructions_per_env_step"] = l0["smsp__inst_executed.sum [inst]"] * 32 / particles / 32
        summary["thread_instructions_per_env_step"] = (
            l0["smsp__inst_executed.sum [inst]"] * l0["smsp__thread_inst_executed_per_inst_executed.ratio []"] / particles)
    json.dump(summary, open(out + ".json", "w"), indent=1)
    print(out + ".json written")


if __name__ == "__main__":
    main()
