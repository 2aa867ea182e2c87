"""Executed-instruction mix by opcode for one kernel: python scripts/ncu_mix.py rep kernel_regex"""
import csv, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
mix = collections.Counter()
tot = 0
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    try:
        n = float(r[idx["Instructions Executed"]])
    except Exception:
        continue
    src = r[idx["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    mix[op] += n
    tot += n
print("total warp instructions", tot)
for op, n in mix.most_common(40):
    print(f"{op:12s} {n/tot*100:5.1f}%  {n:.0f}")
