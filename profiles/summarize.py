"""Turns an ncu report (gpurun_out/*.ncu-rep) into the small tracked summaries under profiles/.
usage: python profiles/summarize.py gpurun_out/prof_r1.ncu-rep profiles/r1_full_summary.csv"""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__inst_executed_pipe_fp32.sum", "smsp__inst_executed.sum",
           # SM pipe utilisation (the NMS kernels' roofline is issue rate): fp32 = fma + fmaheavy/lite pipes, alu, fp64
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main(rep, out):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + ["%s [%s]" % (m, units[ix[m]]) for m in METRICS if m in ix])
        for d in data:
            w.writerow([d[ix["Kernel Name"]].split("(")[0]] + [d[ix[m]] for m in METRICS if m in ix])
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
