"""Read an `ncu --page source --csv` export: total stall-reason histogram and the hottest SASS instructions.
Usage: python tools/ncu_stalls.py gpurun_out/src_users_attention_tc.csv [top_n]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
print(rows[0][1] if len(rows[0]) > 1 else rows[0])
hdr = rows[1]
body = [r for r in rows[2:] if len(r) >= len(hdr) - 2]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stall_cols}
samples_all = 0
inst_total = 0
for r in body:
    for h in stall_cols:
        try:
            tot[h] += int(r[ci[h]])
        except ValueError:
            pass
    try:
        samples_all += int(r[ci["# Samples"]])
        inst_total += int(r[ci["Instructions Executed"]])
    except ValueError:
        pass
print(f"samples {samples_all}  warp-instructions executed {inst_total}  SASS lines {len(body)}")
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:28s} {v:8d}  {100.0 * v / max(samples_all, 1):5.1f}%")
print("hottest instructions (samples, executed, SASS):")
def s(r):
    try:
        return int(r[ci["# Samples"]])
    except ValueError:
        return 0
for r in sorted(body, key=s, reverse=True)[:top]:
    reasons = sorted(((int(r[ci[h]]) if r[ci[h]].isdigit() else 0, h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"  {s(r):6d} {r[ci['Instructions Executed']]:>9s}  {r[ci['Source']].strip()[:90]:90s} {reasons}")
