"""The committed measurement artefacts must agree with each other (no GPU): bench.py copies roofline.traffic from
profiles/r1_traffic.json, and that figure must be what the committed ncu launch list says about the A^7 multiply."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def test_traffic_json_is_the_sum_over_the_committed_launch_list(tmp_path):
    out = tmp_path / "traffic.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"),
                        os.path.join(PROF, "r1_s40_launches_chain30.csv"), str(out), "221545092"],
                       capture_output=True, text=True, check=True)
    fresh, committed = json.load(open(out)), json.load(open(os.path.join(PROF, "r1_traffic.json")))
    assert fresh["traffic"] == committed["traffic"] == committed["dram_bytes_read"] + committed["dram_bytes_write"]
    assert committed["algorithmic_bytes"] == 221545092
    # the step cut out of the list is one full chain: six multiplies, each opened by a pre-pass and closed by a compaction
    assert r.stdout.count("\nA^") == 6
    for p in range(2, 8):
        block = r.stdout.split(f"\nA^{p}:")[1].split("\nA^")[0]
        lines = [l for l in block.splitlines()[1:] if l.strip()]
        assert lines[0].lstrip().startswith("k_prepass") and "k_compact_rows" in block and "k_scan_rowptr" in block


def test_bench_lines_carry_the_contract_keys():
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"}
    for n in (1, 2, 4, 8):
        d = json.load(open(os.path.join(PROF, f"r1_bench_n{n}.json")))
        assert need <= set(d), sorted(need - set(d))
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["gpu_launches"] > 0 and d["config"]["workload"]
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        rf = d["roofline"]
        assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["bound"] == "hbm"
    d1 = json.load(open(os.path.join(PROF, "r1_bench_n1.json")))
    assert d1["cpu_baseline"]["kind"] == "port" and d1["e2e"]["d2h_bytes_per_step"] > 0 and d1["e2e"]["h2d_bytes_per_step"] > 0


# ------------------------------------------------------------------ round 2
def _launches(path):
    import csv
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    idx = {h: i for i, h in enumerate(rows[0])}
    out = {}
    for r in rows[1:]:
        d = out.setdefault(int(r[idx["ID"]]), {"name": r[idx["Kernel Name"]]})
        d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
    return [out[k] for k in sorted(out)]


def test_round2_launch_list_traffic_and_bench_agree():
    """The last timed step of the committed ncu launch list is 17 launches of the engine's own kernels, its A^7 multiply is
    ONE launch whose DRAM bytes match profiles/r2_traffic.json within the run-to-run spread, and the kernel's share of the
    serialised step agrees with the share the bench line measures live."""
    L = _launches(os.path.join(PROF, "r2_launches_chain30.csv"))
    last_flush = max(i for i, d in enumerate(L) if "FillFunctor" in d["name"])
    step = L[last_flush + 1:]
    assert len(step) == 17 and all(d["name"].startswith("void k_") for d in step)
    a7 = step[-1]
    assert "k_lm" in a7["name"]                                             # the operands commute: evaluated as A x A^6 (leftmul.cu)
    tj = json.load(open(os.path.join(PROF, "r2_traffic.json")))
    listed = a7["dram__bytes_read.sum"] + a7["dram__bytes_write.sum"]
    assert abs(listed - tj["traffic"]) / tj["traffic"] < 0.05
    assert tj["traffic"] < 1.2 * 221545092                                  # C is written once: no scratch CSR, no compaction
    share_ncu = a7["gpu__time_duration.sum"] / sum(d["gpu__time_duration.sum"] for d in step)
    b = json.load(open(os.path.join(PROF, "r2_bench_n1.json")))
    share_live = b["per_power"][-1]["ms"] / b["ms_per_step"]
    assert abs(share_ncu - share_live) < 0.05
    assert b["parity"].startswith("A^2..A^7 of this run bit-identical") and b["per_power"][-1]["launches"] == 1


def test_round2_scale_lines_carry_the_contract_keys():
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "per_rank"}
    for kind, ns, scaling in (("weak30", (1, 2, 4, 8), "weak"), ("strong200", (1, 2, 4, 8), "strong"), ("rmat22", (1, 2, 4, 8), "strong"), ("rmat24", (4, 8), "strong")):
        for n in ns:
            d = json.load(open(os.path.join(PROF, f"r2_{kind}_n{n}.json")))
            assert need <= set(d), (kind, n, sorted(need - set(d)))
            assert d["n_gpus"] == n and d["scaling"] == scaling and d["gpu_launches"] > 0 and len(d["per_rank"]) == n
            rf = d["roofline"]
            assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["bound"] == "hbm"
