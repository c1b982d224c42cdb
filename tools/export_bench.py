"""A/B of the two multi-GPU assembly schemes on the weak-scaling mesh of bench.py (n^3 hex8 row elements per rank):
  owner-computes : every column element of the rank, rows of ghost nodes dropped (the product's default path)
  row export     : row elements only, off-rank rows exported over NCCL and added by their owners (FE_MATRIX mode)
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/export_bench.py [n] [steps]
Rank 0 prints one JSON line per scheme (device time of `steps` b200s_evaluate_async calls, max over the ranks)."""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b200 = importlib.import_module("4c_b200")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)) + "/..")
import bench  # noqa: E402  (global_box: the box of N ranks with n^3 elements each)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
capi = b200.capi
box, dims = bench.global_box(n, world)
mesh_of = lambda q: b200.partition.box_partition(b200.mesh.HEX8, box, q, world, dims)
par = [200.0, 0.25]


def timed(ctx, label, n_ele, extra):
    for _ in range(3):
        ctx.evaluate_async(capi.ACTION_NLNSTIFF)
    ctx.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(ctx.device_ptrs()["stream"])
    with torch.cuda.stream(st):
        ev0.record()
        for _ in range(steps):
            ctx.evaluate_async(capi.ACTION_NLNSTIFF)
        ev1.record()
    ctx.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1) / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tot = torch.tensor([float(n_ele)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    if rank == 0:
        print(json.dumps({"scheme": label, "n_gpus": world, "box": list(box), "ms_per_step": float(ms), "row_elements": int(tot),
                          "elements_per_s": float(tot) / (float(ms) * 1e-3), **extra}), flush=True)


t0 = time.time()
rm = mesh_of(rank)
u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / max(box))
asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, par, device=local)
uid = [asm.ctx.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
asm.ctx.comm_init_nccl(uid[0])
asm.set_state("displacement", u[:rm.n_row_nodes].ravel())
timed(asm.ctx, "owner-computes (column elements)", rm.n_row_eles, {"elements_evaluated_rank0": int(rm.n_col_eles)})
K_ref = asm.ctx.get_values()
f_ref = asm.ctx.get_f_rows(0, 3 * rm.n_row_nodes)
asm.close()

view, rowptr, colind, plan = b200.fe_export.fe_plan(mesh_of, rank, world)
t_plan = time.time() - t0
exp = b200.RowExportAssembler(view, rowptr, colind, plan, capi.MAT_COUP_NEOHOOKE, par, device=local)
uid = [exp.ctx.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
exp.ctx.comm_init_nccl(uid[0])
exp.set_state("displacement", u[:rm.n_row_nodes].ravel())
sent = sum(c for _, c in plan.send_val)
timed(exp.ctx, "row elements + export of off-rank rows", rm.n_row_eles,
      {"elements_evaluated_rank0": int(len(view.ele_ids)), "export_values_sent_rank0": int(sent), "plan_s": round(t_plan, 1)})
K, f = exp.own_rows()
scale = np.abs(K_ref).max()
ok = np.abs(K - K_ref).max() <= 1e-12 * scale and np.abs(f - f_ref).max() <= 1e-12 * np.abs(f_ref).max()
flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"check": "own rows of both schemes agree to 1e-12 of the largest entry on every rank", "ok": bool(flag.item() == 1.0)}), flush=True)
exp.close()
dist.barrier()
dist.destroy_process_group()
