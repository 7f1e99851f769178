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
