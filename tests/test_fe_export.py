"""Row-element evaluation + export of the off-rank rows (FE_MATRIX mode, 4C_linalg_sparsematrix.cpp:598-630,705-710):
host-side plan of 4c_b200.fe_export checked on the CPU with the oracle doing the element work -- every rank evaluates
its row elements into owned + export rows, the export rows are "sent" and added at the positions of the plan, and the
result equals the standard owner-computes assembly over the column elements."""
import importlib

import numpy as np
import pytest

from oracle import oracle as orc

b200 = importlib.import_module("4c_b200")
fe = importlib.import_module("4c_b200.fe_export")


def _cases():
    yield "hex8", orc.HEX8, lambda r, w: b200.partition.box_partition(b200.mesh.HEX8, (5, 4, 3), r, w)
    X, conn = b200.mesh.unstructured_tet10(3)
    owner = (X[:, 0] > 0.5).astype(int) + 2 * (X[:, 1] > 0.5).astype(int)
    yield "tet10", orc.TET10, lambda r, w: b200.partition.generic_partition(b200.mesh.TET10, X, conn, owner % w, r, w)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case", ["hex8", "tet10"])
def test_export_plan_reproduces_the_owner_computes_assembly(case, world):
    name, ct, mesh_of = next(c for c in _cases() if c[0] == case)
    par = [200.0, 0.25]
    graphs = {}
    ranks = []
    for r in range(world):
        view, rowptr, colind, plan = fe.fe_plan(lambda q: mesh_of(q, world), r, world, graphs)
        rm = view.rm
        # the graph of the owned rows is the reference's (bit-exact with the oracle's FillComplete restatement)
        o_rowptr, o_colind = orc.build_graph(ct, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
        assert np.array_equal(rowptr[:plan.n_rows_owned + 1], o_rowptr) and np.array_equal(colind[:plan.nnz_owned], o_colind)
        u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.2)
        # row elements only, owned + export rows
        K, f, nbad = orc.evaluate_assemble(ct, orc.MAT_COUP_NEOHOOKE, par, view.conn, rm.X.ravel(), u.ravel(), view.node_row_lid_ext,
                                           rm.node_col_lid, view.n_row_nodes_ext, rowptr, colind)
        assert nbad == 0
        # the standard path: every column element, owned rows only
        rK, rf, _ = orc.evaluate_assemble(ct, orc.MAT_COUP_NEOHOOKE, par, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid,
                                          rm.node_col_lid, rm.n_row_nodes, o_rowptr, o_colind)
        ranks.append((plan, K, f, rK, rf, o_rowptr))
    # every element is a row element of exactly one rank
    assert sum(len(graphs[r][0].ele_ids) for r in range(world)) == len(np.unique(np.concatenate([graphs[r][0].rm.ele_gid for r in range(world)])))
    n_sent = 0
    for r in range(world):
        plan, K, f, rK, rf, o_rowptr = ranks[r]
        Ko, fo = K[:plan.nnz_owned].copy(), f[:plan.n_rows_owned].copy()
        for p, pos, fpos in zip(plan.peers, plan.recv_pos, plan.recv_fpos):
            pp = ranks[p][0]
            i = pp.peers.index(r)
            (v0, nv), (f0, nf) = pp.send_val[i], pp.send_f[i]
            assert nv == len(pos) and nf == len(fpos)
            np.add.at(Ko, pos, ranks[p][1][v0:v0 + nv])
            np.add.at(fo, fpos, ranks[p][2][f0:f0 + nf])
            n_sent += nv
        ok, wK, wf = b200.parity_report(o_rowptr, Ko, rK, fo, rf, 1e-12)
        assert ok, (r, wK, wf)
    assert n_sent > 0
