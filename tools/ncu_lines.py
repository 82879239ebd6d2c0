"""Top CUDA source lines of one kernel by stall samples (from `ncu --page source --csv --print-source sass,cuda`)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None; cur_file = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Name": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "-") and r[0].isdigit():
        col = {h: i for i, h in enumerate(hdr)}
        # columns named 'Source' twice: first is the CUDA line
        def f(k):
            try: return float(r[col[k]])
            except Exception: return 0.0
        stalls = {h[6:]: f(h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
        out.append((f("# Samples"), f("Instructions Executed"), cur_file, int(r[0]), r[1].strip(), stalls))
tot = sum(o[0] for o in out); toti = sum(o[1] for o in out)
print(f"total samples {tot:.0f} instructions {toti:.0f}")
for s, i, fl, ln, src, st in sorted(out, key=lambda o: -o[0])[:n]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*s/max(tot,1):5.1f}% smp {100*i/max(toti,1):5.1f}% ins  {fl}:{ln:<4d} {','.join(f'{k}={v:.0f}' for k,v in top if v>0):40s} {src[:100]}")
