"""Debug helper: CG history on a partitioned mesh vs the same mesh on one rank."""
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b200 = importlib.import_module("4c_b200")
capi = b200.capi
sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
from mgpu_worker import _field

def run(rank, world, local, n, comm):
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, rank, world)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / max(n))
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], device=local)
    if comm:
        uid = [asm.ctx.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        asm.ctx.comm_init_nccl(uid[0])
    nr = rm.n_row_nodes
    asm.set_state("displacement", 0.1 * u[:nr].ravel())
    asm.evaluate_resident("struct_calc_nlnstiff")
    toggle = np.repeat((rm.X[:nr, 0] < 1e-12).astype(np.float64), 3)
    asm.ctx.apply_dirichlet(toggle, True)
    b = _field(rm.node_gid[:nr], 3.0).ravel()
    b[toggle == 1.0] = 0.0
    out = []
    for maxit in (1, 2, 3, 4, 8):
        try:
            x, it, rel = asm.ctx.solve_cg(b, rtol=1e-30, maxit=maxit)
            y = asm.ctx.multiply(x)
            out.append((maxit, rel, float(np.abs(x).max()), float(np.linalg.norm(b - y))))
        except Exception as ex:
            out.append((maxit, str(ex)))
    asm.close()
    return out

if __name__ == "__main__":
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = (12, 7, 5)
    part = run(rank, world, local, n, True)
    print("PART", rank, part, flush=True)
    dist.barrier()
    if rank == 0:
        print("SERIAL", run(0, 1, local, n, False), flush=True)
    dist.barrier()
    dist.destroy_process_group()
