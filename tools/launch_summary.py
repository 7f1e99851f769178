"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of bench.py:
cut out the last complete A^2..A^7 chain that follows an L2 flush (a timed step), print the kernels of every multiply with
their share of the step, and write the A^7 multiply's DRAM bytes to a JSON file (bench.py's roofline.traffic).

  python tools/launch_summary.py gpurun_out/launches.csv [profiles/r1_traffic.json] [algorithmic_bytes_of_A7]
"""
import collections, csv, json, sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    d = L.setdefault(r[idx["ID"]], {"name": r[idx["Kernel Name"]].replace("void ", ""), "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]})
    v = float(r[idx["Metric Value"]].replace(",", "")); u = r[idx["Metric Unit"]]
    if u == "ns": v /= 1e3
    if u == "ms": v *= 1e3
    if u == "Kbyte": v *= 1e3
    if u == "Mbyte": v *= 1e6
    if u == "Gbyte": v *= 1e9
    d[r[idx["Metric Name"]]] = v
ids = list(L)
is_flush = lambda d: "elementwise" in d["name"] or "fill" in d["name"].lower()
# a chain = the engine kernels between two flush kernels; a multiply ends with its placement kernel (k_compact_rows: binned
# pipeline) or IS one launch (k_rw_fused, k_lm, k_dn); helper kernels (descriptors, packs, bounds) belong to the multiply they precede
ENDS = ("k_compact_rows", "k_rw_fused", "k_lm", "k_dn")
flushes = [i for i, k in enumerate(ids) if is_flush(L[k])]
best = None
for f in flushes:
    j = f + 1; run = []; done = 0
    while j < len(ids) and not is_flush(L[ids[j]]):
        # (format checks and per-operand caches of an upload that happen to follow the flush are not part of a multiply)
        if run or not L[ids[j]]["name"].startswith(("k_value_stats", "k_rows_sorted", "k_rowptr_stats", "k_build_desc", "k_build_pack", "k_cspan_bounds", "k_fz_row_span")): run.append(ids[j])
        if L[ids[j]]["name"].startswith(ENDS): done += 1
        j += 1
    if done == 6: best = run
if best is None:
    raise SystemExit("no complete chain (6 multiplies) after an L2 flush found")
step_us = sum(L[k]["gpu__time_duration.sum"] for k in best)
print(f"last timed step: {len(best)} kernels, {step_us:.1f} us serialised by the profiler (compare shares, not absolutes)")
mult, cur = [], []
for k in best:
    cur.append(k)
    if L[k]["name"].startswith(ENDS):
        mult.append(cur); cur = []
if cur: mult[-1].extend(cur)
for p, m in enumerate(mult, start=2):
    t = sum(L[k]["gpu__time_duration.sum"] for k in m)
    rd = sum(L[k].get("dram__bytes_read.sum", 0) for k in m); wr = sum(L[k].get("dram__bytes_write.sum", 0) for k in m)
    print(f"\nA^{p}: {len(m)} kernels, {t:.1f} us ({100 * t / step_us:.1f}% of the step), dram read {rd / 1e6:.2f} MB + write {wr / 1e6:.2f} MB = {(rd + wr) / 1e6:.2f} MB")
    for k in m:
        d = L[k]
        nm = d["name"].split("(")[0][:44]
        print(f"  {nm:44s} {d['grid']:>14s} {d['block']:>12s} {d['gpu__time_duration.sum']:8.1f} us  rd {d.get('dram__bytes_read.sum', 0) / 1e6:8.2f} MB  wr {d.get('dram__bytes_write.sum', 0) / 1e6:8.2f} MB"
              f"  {100 * d['gpu__time_duration.sum'] / t:5.1f}% of the multiply")
if len(sys.argv) > 2:
    m = mult[-1]
    rd = sum(L[k].get("dram__bytes_read.sum", 0) for k in m); wr = sum(L[k].get("dram__bytes_write.sum", 0) for k in m)
    out = {"workload": "A^7 = A^6 x A, 30^3 torus chain, u64, scratch mode", "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
           "source": f"{path} (ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the multiply's kernels, --cache-control none)"}
    if len(sys.argv) > 3: out["algorithmic_bytes"] = int(sys.argv[3])
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(f"\nwrote {sys.argv[2]}: traffic {rd + wr:.0f} B")
