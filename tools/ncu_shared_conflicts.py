"""Per-opcode shared-memory wavefronts (actual vs ideal) and global sectors of every kernel in an .ncu-rep
captured with --set full --import-source on: finds bank conflicts that the summary metrics hide.
usage: python tools/ncu_shared_conflicts.py file.ncu-rep"""
import collections, csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
def f(x):
    try: return float(x)
    except ValueError: return 0.0
kern, h, data = None, None, []
def flush():
    if not data: return
    agg, ideal, cnt = collections.Counter(), collections.Counter(), collections.Counter()
    for d in data:
        w = d["Source"].split()
        op = (w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "?"))
        agg[op] += f(d["L1 Wavefronts Shared"]); ideal[op] += f(d["L1 Wavefronts Shared Ideal"]); cnt[op] += f(d["Instructions Executed"])
    tot, ti = sum(agg.values()), sum(ideal.values())
    print(f"== {kern[:90]}\n   shared wavefronts {tot:.3g} (ideal {ti:.3g}, x{tot / ti if ti else 0:.2f})")
    for k, v in agg.most_common(6):
        if v: print(f"   {k:24s} {v:12.0f}  ideal {ideal[k]:12.0f}  instr {cnt[k]:12.0f}  x{v / ideal[k] if ideal[k] else 0:.2f}")
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        flush(); kern, h, data = r[1], None, []
    elif "Source" in r and "L1 Wavefronts Shared" in r:
        h = r
    elif h and len(r) == len(h):
        data.append(dict(zip(h, r)))
flush()
