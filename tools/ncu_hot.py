"""Summarise an ncu --page source --csv dump (one kernel): hottest SASS regions by instructions executed and stall samples."""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in data); tot_samp = sum(f(r, "# Samples") for r in data)
print(f"kernel: {rows[0][1][:80]}  SASS lines={len(data)} inst={tot_inst:.0f} samples={tot_samp:.0f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("stall samples:", ", ".join(f"{k[6:]}={v:.0f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print("top SASS by samples:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
    top = max(stalls, key=lambda s: f(r, s))
    print(f"  {r[idx['Address']][-5:]} samp={f(r,'# Samples'):6.0f} inst={f(r,'Instructions Executed'):9.0f} thr={f(r,'Avg. Threads Executed'):4.1f} {top[6:]:12s} {r[idx['Source']][:90]}")
