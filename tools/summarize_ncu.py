"""Writes a markdown summary of one `ncu --set full` report (first captured launch) into profiles/.
usage: python tools/summarize_ncu.py gpurun_out/x.ncu-rep profiles/x.md "<command that was profiled>" "<reading>" """
import csv
import subprocess
import sys

KEEP = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]


def main(rep, dst, command, reading):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    stalls = []
    for i, h in enumerate(hdr):
        if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
            try:
                stalls.append((float(vals[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    with open(dst, "w") as f:
        f.write(f"# ncu --set full: `{vals[hdr.index('Kernel Name')][:110]}`\n\ncommand: `{command}`\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"| {k} | {vals[i]} | {units[i]} |\n")
        f.write("\nstall reasons (warps stalled per issue slot): " + ", ".join(f"{h} {v:.2f}" for v, h in sorted(stalls, reverse=True)[:6]) + "\n")
        f.write(f"\n{reading}\n")
    def val(k):
        i = hdr.index(k)
        return float(vals[i]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[i]]
    print(dst, "dram bytes per launch:", int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")), "time", vals[hdr.index("gpu__time_duration.sum")], units[hdr.index("gpu__time_duration.sum")])
    return int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))


if __name__ == "__main__":
    main(*sys.argv[1:5])
