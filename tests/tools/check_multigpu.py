"""torchrun --nproc-per-node 2 tests/tools/check_multigpu.py
Row-sharded run_inference (observation sweeps on shards + replicated latent sweeps after an
all-gather of the per-row state) must reproduce the single-GPU run exactly."""
import os, sys; sys.path.insert(0, '.')  # run from the repo root
import numpy as np, torch, torch.distributed as dist
from pclean_b200.host_fixture import model as M
from pclean_b200.host_fixture.experiments import load_experiment
from pclean_b200.engine import Engine, load_trace_from_snapshot
from oracle import Oracle, export_snapshot

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", lr))
name = sys.argv[1] if len(sys.argv) > 1 else "hospital"
cfg = M.InferenceConfig(2, 4)
model, query, dirty, clean, ir, obs = load_experiment(name, max_rows=4000 if name == "rents" else None)
o = Oracle(ir, M.InferenceConfig(1, 2, use_mh_instead_of_pg=True), seed=5); o.load_observations(obs); o.initialize_trace()
snap = export_snapshot(o, ir, model, query.cls)
n = obs.n_rows; cls = ir.class_index[query.cls]
def make(shard):
    e = Engine(ir, cfg, device=lr); e.load_observations(obs); load_trace_from_snapshot(e, ir, model, query.cls, snap)
    if shard:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0: uid.copy_(torch.tensor(list(Engine.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        e.set_row_shard(cls, (n * rank) // world, (n * (rank + 1)) // world)
        e.nccl_init(bytes(uid.cpu().numpy().tolist()), rank, world)
    return e
es = make(True); st = es.run_inference(9)
e1 = make(False); s1 = e1.run_inference(9)
r0, r1 = (n * rank) // world, (n * (rank + 1)) // world
fks = [v for v, nd in enumerate(model.classes[query.cls].nodes) if isinstance(nd, M.ForeignKeyNode)]
same = all((es.download_assignment(cls, v, n)[r0:r1] == e1.download_assignment(cls, v, n)[r0:r1]).all() for v in fks)
tabs = all(es.table_size(ir.class_index[c]) == e1.table_size(ir.class_index[c]) for c in model.class_order[:-1])
for c in model.class_order[:-1]:
    a = es.download_table(ir.class_index[c]); b = e1.download_table(ir.class_index[c])
    tabs = tabs and (a[1] == b[1]).all()
flag = torch.tensor([int(same and tabs)], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0: print("MULTIGPU", name, "identical to single GPU:", bool(flag.item()), "new_rows", st["new_rows"], s1["new_rows"], "changed", st["changed_rows"], s1["changed_rows"])
dist.destroy_process_group()
