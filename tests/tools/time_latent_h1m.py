"""Per-class device time of one full pgibbs_sweep! on the H1M workload (which latent class dominates?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
from pclean_b200.host_fixture import model as M
from pclean_b200.engine import Engine, load_trace_from_snapshot
from pclean_b200.host_fixture.synth import build_synthetic_hospital
model, query, dirty, truth, ir, obs, snap = build_synthetic_hospital(rows, 20260924)
e = Engine(ir, M.InferenceConfig(1, 20))
e.load_observations(obs)
load_trace_from_snapshot(e, ir, model, query.cls, snap)
for opt in sys.argv[2:]:
    k, v = opt.split("="); e.set_option(k, int(v))
e.sweep(ir.class_index[query.cls], 1, 1)
for sweep in (2, 3):
    for name in model.class_order:
        c = ir.class_index[name]
        t0 = time.time()
        e.debug_counters()
        st = e.sweep(c, 1, sweep)
        print("   pruned-path counters", e.debug_counters()[:16])
        print(sweep, name, "wall_ms %.1f" % (1000 * (time.time() - t0)), "device_ms %.1f" % st["total_ms"], "kernel_ms %.1f" % st["kernel_ms"],
              "changed", st["changed_rows"], "new", st["new_rows"], "launches", st["launches"], flush=True)
