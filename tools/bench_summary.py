"""One-screen summary of a bench.py JSON line: python tools/bench_summary.py file.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(f"headline {d['value']/1e3:.1f} GS/s  {d['ms_per_step']:.4f} ms  whole_step {r['whole_step_frac']:.4f}  dominant {r['kernel']} frac {r['frac']:.4f}")
print("  kernels", {k: round(v, 4) for k, v in r["kernel_ms_per_step"].items()}, "parity", (d.get("parity") or {}).get("parity_checked"))
e = d.get("e2e")
if e: print(f"  e2e {e['value']/1e3:.2f} GS/s  h2d {e.get('h2d_gbs', 0):.1f} GB/s  ceiling {e.get('copy_ceiling_gbs', 0):.1f}  frac {e.get('frac_of_copy_ceiling')}")
if d.get("cpu_baseline"): print("  cpu", round(d["cpu_baseline"]["value"], 1), "MS/s on", d["cpu_baseline"]["cores"], "cores", d["cpu_baseline"]["kind"])
for k, v in (d.get("other_configs") or {}).items():
    if "error" in v:
        print(" ", k, "ERROR", v["error"]); continue
    if "achieved_gbs" in v:
        print(f"  {k:26s} {v['value']/1e3:8.1f} GS/s {v['ms_per_step']:8.3f} ms  {v['achieved_gbs']:.0f} GB/s = {v['frac_of_hbm_peak']:.3f} of HBM peak  parity {(v.get('parity') or {}).get('parity_checked')}/{(v.get('parity') or {}).get('max_abs_lsb')}"); continue
    c = v.get("fp32_pipe_ceiling", {})
    print(f"  {k:26s} {v['value']/1e3:8.1f} GS/s {v['ms_per_step']:8.3f} ms  hbm {v['whole_step_frac']:.4f}  {v['variant']:11s} parity {v['parity']['parity_checked']}/{v['parity']['max_abs_lsb']}"
          + (f"  fp32-ceiling frac {c['frac_of_ceiling']:.3f}" if c else ""))
    print("      ", {a: round(b, 3) for a, b in v["kernel_ms_per_step"].items()})
