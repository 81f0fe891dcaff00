"""Turn an .ncu-rep (brought back under gpurun_out/) into a small committed summary under profiles/."""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        rec = {"kernel": vals[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                rec[k] = "%s %s" % (vals[i], units[i])
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(vals[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        total = sum(v for v, _ in stalls) or 1.0
        rec["stall_samples_pct"] = {name: round(100 * v / total, 1) for v, name in sorted(stalls, reverse=True)[:8]}
        kernels.append(rec)
    with open(out, "w") as f:
        json.dump({"source": rep, "command": "ncu --set full --clock-control none --import-source on", "kernels": kernels}, f, indent=1)
    print(out, len(kernels), "kernel(s)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
