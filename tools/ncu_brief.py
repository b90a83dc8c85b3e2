"""Brief view of an .ncu-rep: key throughput metrics, stall mix and the hottest SASS lines.
usage: python tools/ncu_brief.py file.ncu-rep [n_lines]"""
import csv, subprocess, sys
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
units = dict(zip(hdr, rows[1]))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:80])
    for k in KEYS:
        if k in d: print(f"  {k} = {d[k]} {units.get(k, '')}")
    st = []
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            try: st.append((float(d[k]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError: pass
    print("  stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
for h in hi[:1]:
    hdr = rows[h]; ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = []
    for n, r in enumerate(rows[h + 1:]):
        try: data.append((int(r[isamp]), int(r[iex]), n, r[ia][:90]))
        except (ValueError, IndexError): pass
    tot = sum(x[0] for x in data) or 1
    print("  total samples", tot, "warp instructions", sum(x[1] for x in data))
    for x in sorted(data, reverse=True)[:nl]: print(f"  {100 * x[0] / tot:5.1f}%  exec {x[1]:>9}  #{x[2]:<5} {x[3]}")
