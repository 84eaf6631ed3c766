"""The committed measurement artefacts that bench.py reads are consistent with the ncu exports they came from
(profiles/COMMANDS.md): the DRAM bytes per k_block launch in kblock_traffic_r2.json — bench.py's
`roofline.traffic` — are the sums of dram__bytes_read/write of profiles/ncu_kblock_r2_final_raw.csv."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def test_traffic_file_matches_the_ncu_export():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), os.path.join(PROF, "ncu_kblock_r2_final_raw.csv")],
                         capture_output=True, text=True, check=True).stdout
    launches = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    assert len(launches) == 2 and all("k_block" in d["kernel"] for d in launches)
    traffic = json.load(open(os.path.join(PROF, "kblock_traffic_r2.json")))
    ctx = json.load(open(os.path.join(PROF, "kblock_ncu_r2.json")))
    for i, d in enumerate(launches):
        key = f"k_block(block={i})"
        assert abs(traffic[key] - d["dram_bytes"]) < 1.0
        assert abs(ctx[key]["dram_bytes"] - d["dram_bytes"]) < 1.0
        assert 0.0 < d["dram_gbs"] < 6478.6            # a measured rate, below the measured HBM copy peak


def test_bench_lines_carry_the_contract_keys():
    for name in ("bench_r2_h1m.json", "bench_r2_h1m_8gpu.json", "bench_r2_r10m_1gpu.json", "bench_r2_reference_arm.json"):
        d = json.loads(open(os.path.join(PROF, name)).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config", "e2e"):
            assert k in d, (name, k)
        assert d["metric"] == "gibbs_sweep_rows_x_particles_per_sec" and d["value"] > 0
        if d.get("impl") != "reference":
            assert d["gpu_launches"] > 0 and d["clocks"]["reasons"] == []
