"""Worker of tests/test_gpu_multirank.py (launched by torch.distributed.run, one rank per GPU):
evaluate + assemble with the NCCL ghost-displacement halo of the C-ABI library and compare the rank's
owned rows with the oracle evaluated on the rank's column elements and the exact column state."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
b200 = importlib.import_module("4c_b200")
from oracle import oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n = tuple(int(x) for x in sys.argv[1].split(","))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    capi = b200.capi
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, rank, world)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / max(n))
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], device=local)
    uid = [asm.ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    asm.ctx.comm_init_nccl(uid[0])
    rowptr, colind = asm.graph()
    o_rowptr, o_colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    assert rowptr.tobytes() == o_rowptr.tobytes() and colind.tobytes() == o_colind.tobytes()
    worst = 0.0
    for it in range(2):                      # second pass: new state, halo buffers reused
        uu = u * (1.0 + 0.25 * it)
        asm.set_state("displacement", uu[:rm.n_row_nodes].ravel())   # owned dofs only: ghosts via NCCL
        K, f, err = asm.evaluate("struct_calc_nlnstiff")
        rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, orc.MAT_COUP_NEOHOOKE, [200.0, 0.25], rm.conn,
                                             rm.X.ravel(), uu.ravel(), rm.node_row_lid, rm.node_col_lid,
                                             rm.n_row_nodes, rowptr, colind, nthreads=4)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-12)
        assert ok and nbad == 0 and err.n_neg_jacobian == 0, (rank, it, wK, wf)
        worst = max(worst, wK, wf)
    t = asm.ctx.timers()
    asm.close()
    dist.barrier()
    print(f"MGPU_OK rank {rank}/{world} rows {3 * rm.n_row_nodes} peers {rm.peers} worst ratio {worst:.3g} "
          f"halo_ms {t['halo']:.3f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
