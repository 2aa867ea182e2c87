"""Top stalled SASS instructions of one kernel in an .ncu-rep: python scripts/ncu_hot.py rep kernel_regex [N]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel instance only
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = []
tot = 0
for i, r in enumerate(rows[2:]):
    if r and r[0] == "Kernel Name":
        break
    try:
        v = float(r[idx["Warp Stall Sampling (All Samples)"]])
    except Exception:
        continue
    data.append((v, i, r))
    tot += v
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for v, i, r in sorted(data, key=lambda t: -t[0])[:N]:
    top = sorted(((float(r[idx[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{v/tot*100:5.1f}% #{i:5d} {r[idx['Source']].strip()[:90]:90s} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}")
