"""Summarise ncu output for profiles/: per-kernel share of a launch list (csv from
`ncu --metrics gpu__time_duration.sum --csv`) and key metrics of a full report
(`ncu -i X.ncu-rep --page raw --csv`)."""
import collections
import csv
import subprocess
import sys


def launch_list(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        k = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        if r[ui] == "us":
            v *= 1e3
        elif r[ui] == "ms":
            v *= 1e6
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["kernel,launches,total_ns,share_pct,avg_ns"]
    for k, (n, t) in agg.items():
        out.append(f"{k},{n},{t:.0f},{100 * t / tot:.2f},{t / n:.0f}")
    return "\n".join(out)


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__icc_request_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum"]


def full_report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append("== " + d["Kernel Name"])
        for k in KEYS:
            if k in d:
                out.append(f"{k} = {d[k]} {units[hdr.index(k)]}")
        st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v.replace(",", "")) for k, v in d.items()
              if "pcsamp_warps_issue_stalled" in k and not k.endswith("not_issued") and v not in ("", "n/a")}
        tot = sum(st.values()) or 1
        out.append("stall samples (% of all): " + ", ".join(f"{k} {100 * v / tot:.1f}" for k, v in
                                                            sorted(st.items(), key=lambda x: -x[1])[:8]))
    return "\n".join(out)


if __name__ == "__main__":
    if sys.argv[1] == "list":
        print(launch_list(sys.argv[2]))
    else:
        print(full_report(sys.argv[2]))
