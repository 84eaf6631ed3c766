"""`bench.py --impl reference` (the CPU restatement timed on the host cores, one process per core)
prints the contract's JSON line; runs without a GPU on a small table."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, PCLEAN_BENCH_PROCS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "1500", "--hospitals", "32",
                          "--steps", "1", "--warmup", "0", "--ref-rows", "2"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "gibbs_sweep_rows_x_particles_per_sec"
    assert line["value"] > 0 and line["unit"] == "rows*particles/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["cores"] == 2 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"]
