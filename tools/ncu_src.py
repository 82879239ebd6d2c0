"""Top SASS instructions of one kernel by stall samples / instruction count / shared-memory conflicts (from `ncu --page source --csv`)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
def f(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
tot_s = sum(f(r, "# Samples") for r in data); tot_i = sum(f(r, "Instructions Executed") for r in data)
print(f"total samples {tot_s:.0f}, instructions {tot_i:.0f}, SASS lines {len(data)}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("stall totals:", ", ".join(f"{k[6:]}={v:.0f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("-- top by samples")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:n]:
    top = max(stalls, key=lambda s: f(r, s))
    print(f"{f(r,'# Samples'):8.0f} {100*f(r,'# Samples')/max(tot_s,1):5.1f}%  exec {f(r,'Instructions Executed'):10.0f}  {top[6:]:12s} conf {f(r,'L1 Conflicts Shared N-Way'):4.1f}  {r[col['Source']].strip()[:90]}")
print("-- top by shared excessive wavefronts")
for r in sorted(data, key=lambda r: -f(r, "L1 Wavefronts Shared Excessive"))[:8]:
    print(f"{f(r,'L1 Wavefronts Shared Excessive'):12.0f} of {f(r,'L1 Wavefronts Shared'):12.0f}  exec {f(r,'Instructions Executed'):10.0f}  {r[col['Source']].strip()[:90]}")
