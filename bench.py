#!/usr/bin/env python
"""bench.py — Gibbs-sweep throughput of the hot path on the synthetic hospital-schema table.

  python bench.py --gpus N --steps K --warmup W            (our CUDA engine, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  (the CPU restatement of the reference path)

One "step" = one `pgibbs_sweep!`-equivalent over the observation class (Record): the K-particle
row moves of every row + the table-update pass.  Workload = BASELINE.json configs[3], the
configuration the north-star target is quoted on: synthetic hospital-schema, 1,000,000 dirty
rows, K = 20 particles (it fits one GPU).  With N > 1 the same table is row-sharded across
ranks (strong scaling) with one NCCL all-reduce of the reference counts per sweep.

Prints ONE JSON line (rank 0).  The oracle (`oracle/`) is only used for the cpu_baseline leg
and for `--impl reference`; the measured product path never touches it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "gibbs_sweep_rows_x_particles_per_sec"
UNIT = "rows*particles/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--particles", type=int, default=20)
    ap.add_argument("--hospitals", type=int, default=4096)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="time budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-seconds", type=float, default=0.0, help="--impl reference: seconds per step (default: 120 s spread over the steps, at most 20 s each)")
    ap.add_argument("--seed", type=int, default=20260924)
    return ap.parse_args()


def workload_config(a):
    scale = {}
    if a.hospitals != 4096:
        scale = dict(H=a.hospitals, P=max(4, a.hospitals // 2), C=max(4, a.hospitals // 8))
    return scale


def build_workload(a, log):
    from pclean_b200.synth import build_synthetic_hospital
    t0 = time.time()
    out = build_synthetic_hospital(a.rows, a.seed, **workload_config(a))
    log(f"workload built in {time.time() - t0:.1f}s: rows={a.rows} strings={len(out[4].strings)}")
    return out


_ORACLE_CACHE = {}


def cpu_reference_leg(a, work, seconds, log):
    """Time the CPU restatement of the reference path (single thread: the reference is
    single-threaded) on a bounded prefix of the same table, against the FULL latent tables."""
    from pclean_b200 import model as M
    from oracle import Oracle
    model, query, dirty, truth, ir, obs, snap = work
    cap = min(a.rows, 4096)
    o = _ORACLE_CACHE.get(os.getpid())
    if o is None:                      # one oracle per process, reused by the following steps
        cfg = M.InferenceConfig(1, a.particles)
        o = Oracle(ir, cfg, seed=a.seed)
        o.load_observations(obs)
        t0 = time.time()
        o.install_snapshot(ir, model, query.cls, snap, n_obs_rows=cap, bump_to_full=True)
        log(f"oracle trace installed in {time.time() - t0:.1f}s")
        _ORACLE_CACHE[os.getpid()] = o
    cls = ir.class_index[query.cls]
    o.begin_sweep()
    done, chunk = 0, 2
    t0 = time.perf_counter()
    while done < cap and time.perf_counter() - t0 < seconds:
        o.sweep_class(cls, done, min(cap, done + chunk))
        done = min(cap, done + chunk)
    dt = time.perf_counter() - t0
    return dict(value=done * a.particles / dt, unit=UNIT, cores=1, kind="port",
                sample=f"{done} rows x {a.particles} particles of the same table (full latent tables and option lists) in {dt:.1f}s, "
                       f"single thread (the reference is single-threaded), oracle/pclean_oracle.cpp -O2"), done, dt


_REF_WORK = None


def _ref_worker(i):
    a, work, seconds = _REF_WORK
    _, done, dt = cpu_reference_leg(a, work, seconds, lambda m: None)
    return done, dt


class ClockSampler:
    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t_begin=None, t_end=None):
        """median SM clock / throttle reasons over the samples taken inside [t_begin, t_end] (time.time());
        nvidia-smi needs ~0.5 s to start, so it is launched long before the timed region"""
        import datetime
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        rows = []
        for line in self.f.read().strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]), parts[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if t_begin is None or (t_begin - 0.05 <= r[0] <= t_end + 0.05)]
        if not inside and rows:           # timed region shorter than the sampling period: the samples closest to it (under the same load)
            inside = sorted(rows, key=lambda r: abs(r[0] - (t_end if t_end else r[0])))[:3]
        for ts, a, b, flags in inside:
            sm.append(a); mx.append(b)
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        a.gpus = world

    def log(msg):
        print(f"[bench r{rank}] {msg}", file=sys.stderr, flush=True)

    config = {"workload": f"synthetic hospital-schema {a.rows} rows K={a.particles} (BASELINE.json configs[3])",
              "rows": a.rows, "particles": a.particles, "hospitals": a.hospitals, "typo_rate": 0.05, "seed": a.seed,
              "parallelism": f"row-shard x{a.gpus}" if a.gpus > 1 else "single GPU"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return
        work = build_workload(a, log)
        per_step = a.ref_seconds if a.ref_seconds > 0 else max(2.0, min(20.0, 120.0 / max(1, a.steps + a.warmup)))
        # The reference is single-threaded; rows of the observation class are independent given the
        # table snapshot, so the port is run in one process per host core (fork: the workload is
        # shared copy-on-write), each timing the same bounded prefix, and the throughputs add up.
        import multiprocessing as mp
        procs = max(1, min(int(os.environ.get("PCLEAN_BENCH_PROCS", "0")) or (os.cpu_count() or 1), 64))   # each process holds its own copy of the observations
        global _REF_WORK
        _REF_WORK = (a, work, per_step)
        vals = []
        with mp.get_context("fork").Pool(procs) as pool:
            for s in range(a.warmup + a.steps):
                res = pool.map(_ref_worker, range(procs))
                if s >= a.warmup:
                    vals.append((sum(d for d, _ in res), max(t for _, t in res)))
        rows = sum(d for d, _ in vals); secs = sum(t for _, t in vals)
        value = rows * a.particles / secs
        cb = dict(value=value, unit=UNIT, cores=procs, kind="port",
                  sample=f"{rows} rows x {a.particles} particles over {a.steps} steps (prefix of the same table, full latent tables), "
                         f"{procs} processes x 1 thread (oracle/pclean_oracle.cpp -O2; the reference itself is single-threaded)")
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": 1000.0 * secs / max(1, a.steps), "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": cb,
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm (CUDA)
    import numpy as np
    import torch
    import torch.distributed as dist
    from pclean_b200 import model as M
    from pclean_b200.engine import Engine, load_trace_from_snapshot

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None      # samples are filtered to the timed region afterwards
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    work = build_workload(a, log)
    model, query, dirty, truth, ir, obs, snap = work
    cfg = M.InferenceConfig(1, a.particles)
    t0 = time.time()
    e = Engine(ir, cfg, device=local_rank)
    e.load_observations(obs)
    load_trace_from_snapshot(e, ir, model, query.cls, snap)
    cls = ir.class_index[query.cls]
    nb = len(model.classes[query.cls].blocks)
    r0, r1 = (a.rows * rank) // world, (a.rows * (rank + 1)) // world
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.tensor(list(Engine.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        e.set_row_shard(cls, r0, r1)
        e.nccl_init(bytes(uid.cpu().tolist()), rank, world)
    st = e.sweep(cls, a.seed, 1)      # first sweep also builds every distance matrix (setup, untimed)
    torch.cuda.synchronize()
    log(f"engine ready in {time.time() - t0:.1f}s; matrices {e.matrix_bytes() / 2**30:.2f} GiB; first sweep {st}")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from the warm-up steps on (nvidia-smi needs ~0.3 s before its first
    # sample; the GPU is under the same load during warm-up and the timed steps)
    sweep_idx = 2
    for _ in range(max(0, a.warmup - 1)):
        e.sweep(cls, a.seed, sweep_idx); sweep_idx += 1

    # ---- timed region: device-resident inputs ("value")
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    wall0 = time.perf_counter()
    kernel_ms = [0.0] * nb
    launches = 0
    tot_ms = 0.0
    stats = []
    for _ in range(a.steps):
        s = e.sweep(cls, a.seed, sweep_idx); sweep_idx += 1
        tot_ms += s["total_ms"]; launches += s["launches"]; stats.append(s)
        for b in range(nb):
            kernel_ms[b] += e.block_metrics(b)["kernel_ms"]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop(t_begin, time.time()) if sampler else None
    # device time of the steps (engine CUDA events on its own stream), max over ranks
    t = torch.tensor([tot_ms / 1000.0, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall_s = float(t[0]), float(t[1])
    value = a.rows * a.particles * a.steps / wall_s

    # ---- end-to-end: host buffers in, results out, every step (one untimed pass first: pinning
    # the host columns is a one-off)
    e.resync_observations()
    e.download_logweights(cls, a.rows)
    barrier()
    w0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(a.steps):
        h2d = e.resync_observations()
        e.sweep(cls, a.seed, sweep_idx); sweep_idx += 1
        k0 = e.download_assignment(cls, model.classes[query.cls].names["hosp"] - 1, a.rows)
        k1 = e.download_assignment(cls, model.classes[query.cls].names["metric"] - 1, a.rows)
        lw = e.download_logweights(cls, a.rows)
        d2h = k0.nbytes // 2 + k1.nbytes // 2 + lw.nbytes       # device side: int32 slots + f64 log-weights
    barrier()
    t = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = a.rows * a.particles * a.steps / float(t[0])

    if rank == 0:
        # roofline of the dominant kernel (the block kernel with the larger device time)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        dom = max(range(nb), key=lambda b: kernel_ms[b])
        bm = e.block_metrics(dom)
        rows_rank = r1 - r0
        alg_bytes = bm["distance_bytes_per_row"] * rows_rank + 12.0 * a.particles * rows_rank
        k_ms = kernel_ms[dom] / a.steps
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "kblock_traffic_r1.json")
        if world == 1 and a.rows == 1000000 and a.particles == 20 and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"k_block(block={dom})")     # ncu capture of this exact workload
        roofline = {"bound": "hbm", "kernel": f"k_block(block={dom})", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                    "note": "algorithmic bytes = 1 B per (enumerated element x likelihood term) per row (shared across the K particles "
                            "of a row, which the reference recomputes per particle) + 12 B per row x particle written",
                    "all_blocks_ms": [x / a.steps for x in kernel_ms]}
        cpu_baseline = None
        if not a.no_cpu_baseline and world == 1:
            cpu_baseline, _, _ = cpu_reference_leg(a, work, a.cpu_seconds, log)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": 1000.0 * wall_s / a.steps, "device_ms_per_step": 1000.0 * dev_s / a.steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": dict(config, l2="inputs larger than L2: distance matrices %.1f GiB" % (e.matrix_bytes() / 2**30)),
               "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
               "gpu_launches": int(launches),
               "sweep": {"new_rows": sum(s["new_rows"] for s in stats), "changed_rows": sum(s["changed_rows"] for s in stats),
                         "dummy_draws": sum(s["dummy_draws"] for s in stats)}}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
