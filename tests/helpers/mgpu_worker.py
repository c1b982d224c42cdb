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


def _field(gid, seed):
    """Deterministic nodal vector field that every rank can evaluate for any global node id."""
    g = np.asarray(gid, np.float64)[:, None]
    d = np.arange(3)[None, :]
    return np.sin(0.37 * g + 1.3 * d + seed) + 0.25 * np.cos(0.011 * g * (d + 1) + seed)


def post_checks(rank, world, rm, asm, u, rowptr, colind):
    """The actions around the hot path on a partitioned mesh: every one of them imports a row-map vector
    to the column map through the NCCL halo (multiply: x, mass: acceleration, energy / prestress:
    displacement), and the CG dot products are all-reduced."""
    import scipy.sparse as sp
    capi = b200.capi
    nr = rm.n_row_nodes
    n_rows = 3 * nr
    # --- multiply: y = K x with x on the row map, ghosts through the halo -----------------------------
    K = asm.ctx.get_values()
    x_col = _field(rm.node_gid, 0.5)
    A = sp.csr_matrix((K, colind, rowptr), shape=(n_rows, 3 * rm.n_col_nodes))
    y = asm.ctx.multiply(x_col[:nr].ravel())
    ref = A @ x_col.ravel()
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max(), (rank, "multiply")
    # --- energy of the row elements, summed over the ranks -----------------------------------------
    uu = u * 1.25
    e_dev = asm.calc_energy()
    e_ref = orc.calc_energy(orc.HEX8, orc.MAT_COUP_NEOHOOKE, [200.0, 0.25], rm.conn, rm.X.ravel(), uu.ravel(),
                            ele_is_row=rm.ele_is_row)
    assert abs(e_dev - e_ref) <= 1e-11 * abs(e_ref), (rank, "energy", e_dev, e_ref)
    tot = torch.tensor([e_dev, e_ref, float(rm.ele_is_row.sum())], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    assert abs(tot[0] - tot[1]) <= 1e-11 * abs(tot[1])
    # --- mass action: acceleration imported through the halo ------------------------------------------
    acc = _field(rm.node_gid, 2.0)
    asm.ctx.set_density(7.8)
    asm.set_state("acceleration", acc[:nr].ravel())
    asm.evaluate("struct_calc_nlnstiffmass")
    M, fi = asm.mass()
    rM, rfi = orc.assemble_mass(orc.HEX8, 7.8, rm.conn, rm.X.ravel(), rm.node_row_lid, rm.node_col_lid, nr, rowptr,
                                colind, acc_col=acc.ravel())
    assert np.abs(M - rM).max() <= 1e-12 * np.abs(rM).max() and np.abs(fi - rfi).max() <= 1e-12 * np.abs(rfi).max()
    # --- follower pressure on the top face: ghost displacements through the halo, owner-computes rows ----
    from tests.helpers import neumann as nm
    case = {"celltype": "hex8", "coords": rm.X.tolist(), "elements": rm.conn.tolist(),
            "neumann_nodes": np.nonzero(rm.X[:, 2] > 1.0 - 1e-9)[0].tolist(),
            "neumann": {"ONOFF": [1, 0, 0], "VAL": [-0.7, 0, 0], "FUNCT": [0, 0, 0]}}
    faces = nm.loaded_faces(case)
    cond = asm.add_surface_neumann("orthopressure", faces, [1, 0, 0], [-0.7, 0, 0])
    asm.set_state("displacement", uu[:nr].ravel())
    asm.evaluate_resident("struct_calc_nlnstiff")
    K0 = asm.ctx.get_values()
    fext = asm.evaluate_neumann(loadlin=True)
    dK = asm.ctx.get_values() - K0
    if len(faces):
        rf, rK = nm.orthopressure_load(case, uu.ravel(), 0.0, faces=faces)
        own = np.nonzero(rm.node_row_lid >= 0)[0]
        src = (3 * own[:, None] + np.arange(3)).ravel()                      # column-node dofs of the owned nodes
        dst = (3 * rm.node_row_lid[own][:, None] + np.arange(3)).ravel()     # the rows they own
        rf_row = np.zeros(n_rows)
        rf_row[dst] = rf[src]
        sel = np.empty(n_rows, np.int64)
        sel[dst] = src
        colperm = (3 * rm.node_col_lid[:, None] + np.arange(3)).ravel()      # column-node dof -> matrix column
        P = sp.csr_matrix((np.ones(len(colperm)), (np.arange(len(colperm)), colperm)), shape=(len(colperm),) * 2)
        ref_dK = nm.add_to_csr_values((rK.tocsr()[sel] @ P).tocsr(), rowptr, colind, np.zeros(len(colind)))
        assert np.abs(fext - rf_row).max() <= 1e-12 * np.abs(rf).max(), (rank, "neumann f")
        assert np.abs(dK - ref_dK).max() <= 1e-10 * np.abs(rK).max() + 1e-15 * np.abs(K0).max(), (rank, "neumann K")
    else:
        assert np.all(fext == 0.0) and np.all(dK == 0.0)
    asm.ctx.clear_neumann()
    # --- Dirichlet + distributed CG on the resident K: residual through the distributed product -------
    # (tangent at a small strain state: CG needs a positive definite operator)
    asm.set_state("displacement", 0.1 * u[:nr].ravel())
    asm.evaluate_resident("struct_calc_nlnstiff")
    toggle = np.repeat((rm.X[:nr, 0] < 1e-12).astype(np.float64), 3)
    asm.ctx.apply_dirichlet(toggle, True)
    b = _field(rm.node_gid[:nr], 3.0).ravel()
    b[toggle == 1.0] = 0.0
    x, iters, rel = asm.ctx.solve_cg(b, rtol=1e-10, maxit=3000)
    r = b - asm.ctx.multiply(x)
    nrm = torch.tensor([float(r @ r), float(b @ b)], dtype=torch.float64, device="cuda")
    dist.all_reduce(nrm)
    true_rel = float(torch.sqrt(nrm[0] / nrm[1]))
    assert rel <= 1e-10 and true_rel <= 1e-8, (rank, iters, rel, true_rel)
    print(f"MGPU_POST rank {rank}: multiply/energy/mass/neumann({len(faces)} faces) ok, CG {iters} iterations, recurrence residual {rel:.2e}, "
          f"true residual {true_rel:.2e}", flush=True)


def export_main(rank, world, local, n):
    """Row-element mode: every rank evaluates its row elements only, the export rows travel over NCCL inside
    b200s_evaluate (ghost displacements in, off-rank rows out) and the rank's own rows equal the owner-computes assembly
    of the oracle over the column elements."""
    capi = b200.capi
    fe = b200.fe_export
    view, rowptr, colind, plan = fe.fe_plan(lambda q: b200.partition.box_partition(b200.mesh.HEX8, n, q, world), rank, world)
    rm = view.rm
    asm = b200.RowExportAssembler(view, rowptr, colind, plan, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], device=local)
    uid = [asm.ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    asm.ctx.comm_init_nccl(uid[0])
    o_rowptr, o_colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    assert np.array_equal(rowptr[:plan.n_rows_owned + 1], o_rowptr) and np.array_equal(colind[:plan.nnz_owned], o_colind)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / max(n))
    worst = 0.0
    for it in range(2):
        uu = u * (1.0 + 0.25 * it)
        asm.set_state("displacement", uu[:rm.n_row_nodes].ravel())
        K, f, err = asm.evaluate("struct_calc_nlnstiff")
        rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, orc.MAT_COUP_NEOHOOKE, [200.0, 0.25], rm.conn, rm.X.ravel(), uu.ravel(),
                                             rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, o_rowptr, o_colind, nthreads=4)
        ok, wK, wf = b200.parity_report(o_rowptr, K, rK, f, rf, 1e-12)
        assert ok and nbad == 0 and err.n_neg_jacobian == 0, (rank, it, wK, wf)
        _, f2, _ = asm.evaluate("struct_calc_internalforce")
        assert np.abs(f2 - rf).max() <= 1e-12 * np.abs(rf).max()
        worst = max(worst, wK, wf)
    sent = sum(c for _, c in plan.send_val)
    asm.close()
    dist.barrier()
    print(f"MGPU_EXPORT_OK rank {rank}/{world} row elements {len(view.ele_ids)} of {rm.n_col_eles} column elements, "
          f"export values sent {sent}, peers {plan.peers}, worst ratio {worst:.3g}", flush=True)
    dist.destroy_process_group()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n = tuple(int(x) for x in sys.argv[1].split(","))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if len(sys.argv) > 2 and sys.argv[2] == "export":
        return export_main(rank, world, local, n)
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
    if len(sys.argv) > 2 and sys.argv[2] == "post":
        post_checks(rank, world, rm, asm, u, rowptr, colind)
    t = asm.ctx.timers()
    asm.close()
    dist.barrier()
    print(f"MGPU_OK rank {rank}/{world} rows {3 * rm.n_row_nodes} peers {rm.peers} worst ratio {worst:.3g} "
          f"halo_ms {t['halo']:.3f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
