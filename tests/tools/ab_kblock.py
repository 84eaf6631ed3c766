"""A/B timings of k_block on the bench workload (H1M): one engine, sweeps under different
PCL_OPT_* masks and k_block geometries.  Prints one JSON line per setting.
  python tests/tools/ab_kblock.py [rows] [out.jsonl]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else None
    from pclean_b200.host_fixture import model as M
    from pclean_b200.engine import Engine, load_trace_from_snapshot
    from pclean_b200.host_fixture.synth import build_synthetic_hospital
    model, query, dirty, truth, ir, obs, snap = build_synthetic_hospital(rows, 20260924)
    e = Engine(ir, M.InferenceConfig(1, 20))
    e.load_observations(obs)
    load_trace_from_snapshot(e, ir, model, query.cls, snap)
    cls = ir.class_index[query.cls]
    t0 = time.time()
    st = e.sweep(cls, 1, 1)
    print("first sweep", time.time() - t0, st, file=sys.stderr)
    sweep = 2

    def measure(tag, n=4):
        nonlocal sweep
        res = []
        for _ in range(n):
            s = e.sweep(cls, 1, sweep); sweep += 1
            res.append((s["total_ms"], e.block_metrics(0)["kernel_ms"], e.block_metrics(1)["kernel_ms"], s["changed_rows"], s["sum_log_ml"]))
        best = min(res)
        line = {"tag": tag, "total_ms": best[0], "block0_ms": best[1], "block1_ms": best[2], "first_total_ms": res[0][0],
                "first_block1_ms": res[0][2], "changed": res[-1][3], "sum_log_ml": res[-1][4]}
        print(json.dumps(line), flush=True)
        if out:
            out.write(json.dumps(line) + "\n"); out.flush()

    ALL = 31
    for kb in (0, 1, 2):
        e.set_option("kb_variant", kb)
        measure(f"kb{kb} opts=31")
    e.set_option("kb_variant", 0)
    for opts, name in ((0, "none"), (15, "all-but-lazynew"), (16, "lazynew"), (ALL - 1, "all-but-progressive"), (ALL - 2, "all-but-pmemo"),
                       (ALL - 8, "all-but-parhint"), (ALL, "all")):
        e.set_option("opts", opts)
        measure(f"kb0 opts={opts} ({name})")
    e.set_option("opts", ALL)
    e.set_option("memo", 0)
    measure("kb0 opts=31 memo=0", n=2)
    e.set_option("memo", 1)
    e.set_option("prune", 0)
    measure("kb0 opts=31 prune=0", n=1)


if __name__ == "__main__":
    main()
