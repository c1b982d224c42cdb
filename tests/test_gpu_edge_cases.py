"""Edge cases of the evaluate + assemble path through the C ABI, against the oracle: empty element blocks, ranks without row
nodes, isolated nodes (empty CSR rows), one-element blocks whose node numbering leaves gaps.  The reference meets all of them
(a processor without elements or without row elements is ordinary in Discretization::evaluate,
src/core/fem/src/discretization/4C_fem_discretization_evaluate.cpp:85-117: the column-element loop simply has nothing to do)."""
import importlib

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
b200 = importlib.import_module("4c_b200")
capi = b200.capi
PAR = [200.0, 0.25]


def _check(rm, u, ocell):
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, PAR)
    rowptr, colind = asm.graph()
    o_rowptr, o_colind = orc.build_graph(ocell, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    assert np.array_equal(rowptr, o_rowptr) and np.array_equal(colind, o_colind)
    assert len(rowptr) == 3 * rm.n_row_nodes + 1 and asm.nnz == rowptr[-1]
    for action, oaction in (("struct_calc_nlnstiff", orc.ACTION_FORCE_STIFF), ("struct_calc_internalforce", orc.ACTION_FORCE)):
        asm.set_state("displacement", u[:rm.n_row_nodes].ravel())
        K, f, err = asm.evaluate(action)
        rK, rf, nbad = orc.evaluate_assemble(ocell, orc.MAT_COUP_NEOHOOKE, PAR, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid,
                                             rm.node_col_lid, rm.n_row_nodes, rowptr, colind, action=oaction)
        assert err.n_neg_jacobian == 0 and nbad == 0
        assert len(f) == 3 * rm.n_row_nodes
        if oaction == orc.ACTION_FORCE_STIFF:
            assert len(K) == rowptr[-1]
            if rowptr[-1]:
                ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-12)
                assert ok, (wK, wf)
        if len(f):
            assert np.abs(f - rf).max() <= 1e-12 * max(np.abs(rf).max(), 1e-300)
    asm.close()
    return rowptr, f


@pytest.mark.parametrize("cell", ["hex8", "tet10", "hex27", "tet4"])
def test_empty_element_block(cell):
    """Nodes, but no elements on this processor: empty rows, zero force, nothing launched that needs an element."""
    ct = {"hex8": b200.mesh.HEX8, "tet10": b200.mesh.TET10, "hex27": b200.mesh.HEX27, "tet4": b200.mesh.TET4}[cell]
    oc = {"hex8": orc.HEX8, "tet10": orc.TET10, "hex27": orc.HEX27, "tet4": orc.TET4}[cell]
    X = np.random.default_rng(0).uniform(0, 1, (6, 3))
    rm = b200.partition.serial_mesh(ct, X, np.zeros((0, orc.num_nodes(oc)), np.int32))
    rowptr, f = _check(rm, 0.01 * X, oc)
    assert rowptr[-1] == 0 and not f.any()


def test_no_nodes_no_elements():
    """A processor that holds nothing at all."""
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, np.zeros((0, 3)), np.zeros((0, 8), np.int32))
    rowptr, f = _check(rm, np.zeros((0, 3)), orc.HEX8)
    assert len(rowptr) == 1 and len(f) == 0


def test_isolated_nodes_have_empty_rows():
    """Nodes no element refers to, in the middle of the numbering: their CSR rows are empty, their force entries zero, the
    rows of the other nodes are what the oracle assembles."""
    rm0 = b200.partition.box_partition(b200.mesh.HEX8, (2, 2, 1), 0, 1)
    n0 = rm0.n_col_nodes
    extra = np.array([[5.0, 5.0, 5.0], [6.0, 5.0, 5.0], [7.0, 7.0, 7.0]])
    # new numbering: isolated nodes at positions 0, 7 and the end
    order = [n0] + list(range(0, 6)) + [n0 + 1] + list(range(6, n0)) + [n0 + 2]
    Xall = np.vstack([rm0.X, extra])[order]
    newid = np.empty(n0 + 3, np.int64)
    newid[order] = np.arange(n0 + 3)
    conn = newid[rm0.conn].astype(np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, Xall, conn)
    u = b200.mesh.displacement_field(Xall, np.arange(len(Xall)), 1.0, 0.1)
    rowptr, f = _check(rm, u, orc.HEX8)
    for k in (0, 7, n0 + 2):
        assert rowptr[3 * k + 3] == rowptr[3 * k] and not f[3 * k:3 * k + 3].any()
    assert rowptr[-1] > 0


def test_rank_with_ghost_elements_only():
    """Column elements, but every node is owned elsewhere (no row nodes): the loop runs over the elements and every
    contribution is dropped (lmrowowner != myrank, 4C_linalg_sparsematrix.cpp:386-447).  One GPU, rank view with nranks = 1
    as in test_hex8_partitioned_rows_match_oracle."""
    rm0 = b200.partition.box_partition(b200.mesh.HEX8, (2, 1, 1), 0, 1)
    rm = b200.partition.RankMesh(rm0.celltype, 0, 1, rm0.X, rm0.conn, rm0.node_gid, np.full(rm0.n_col_nodes, -1, np.int32),
                                 rm0.node_col_lid, 0, 0, rm0.ele_gid)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.1)
    rowptr, f = _check(rm, u, orc.HEX8)
    assert len(rowptr) == 1 and rowptr[-1] == 0 and len(f) == 0


@pytest.mark.parametrize("tech,prestress", [(capi.TECH_NONE, capi.PRESTRESS_NONE), (capi.TECH_FBAR, capi.PRESTRESS_NONE),
                                            (capi.TECH_EAS_MILD, capi.PRESTRESS_NONE), (capi.TECH_EAS_FULL, capi.PRESTRESS_NONE),
                                            (capi.TECH_NONE, capi.PRESTRESS_MULF), (capi.TECH_SHELL_EAS_ANS, capi.PRESTRESS_NONE)])
def test_empty_block_through_every_action(tech, prestress):
    """The actions around the path on a hex8 block without elements (formulations with element / Gauss-point histories
    included): mass + inertia force, stress / strain / energy output, recover, update / reset of the step, prestress update,
    history get / set, device Dirichlet rows -- nothing to do, no error, zero results."""
    X = np.random.default_rng(1).uniform(0, 1, (5, 3))
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, X, np.zeros((0, 8), np.int32))
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, PAR, tech=tech, prestress=prestress, density=1.5)
    assert asm.nnz == 0 and asm.n_rows == 15
    u = 0.01 * X.ravel()
    asm.set_state("displacement", u)
    asm.set_state("acceleration", u)
    for action in ("struct_calc_nlnstiff", "struct_calc_internalforce", "struct_calc_nlnstiffmass", "struct_calc_nlnstifflmass"):
        K, f, err = asm.evaluate(action)
        assert (K is None or len(K) == 0) and not f.any() and err.n_neg_jacobian == 0
    st, en = asm.calc_stress("cauchy", "ea")
    assert st.shape[0] == 0 and en.shape[0] == 0
    assert asm.calc_energy() == 0.0
    if tech in (capi.TECH_EAS_MILD, capi.TECH_EAS_FULL, capi.TECH_SHELL_EAS_ANS):
        asm.recover(np.zeros(15))
        assert asm.ctx.get_element_history().size == 0
    if prestress == capi.PRESTRESS_MULF:
        asm.update_prestress()
    asm.ctx.update_step()
    asm.ctx.reset_step()
    asm.close()


def _with_extra_blocks(rowptr, colind, n_row_nodes, extra_cols_of_node):
    """CSR graph with more node blocks per row node (all three dof rows of a node get the same columns, sorted)."""
    rp, ci = [0], []
    for k in range(n_row_nodes):
        base = set((colind[rowptr[3 * k]:rowptr[3 * k + 1]] // 3).tolist()) | set(extra_cols_of_node(k))
        cols = np.array(sorted(base), np.int64)
        dof = (3 * cols[:, None] + np.arange(3)).ravel()
        for _ in range(3):
            ci.append(dof)
            rp.append(rp[-1] + len(dof))
    return np.asarray(rp, np.int64), np.concatenate(ci).astype(np.int32)


def test_caller_graph_with_more_entries_than_the_block_needs():
    """The matrix of the reference carries the graph of everything that assembles into it (other element blocks, coupling
    conditions): a caller graph that is a superset of the block's own pattern.  The extra entries stay zero, the others are
    what the oracle adds into the same graph (assemble looks positions up per row, 4C_linalg_sparsematrix.cpp:386-447)."""
    rm = b200.partition.box_partition(b200.mesh.HEX8, (4, 3, 2), 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.2)
    rowptr0, colind0 = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    nn = rm.n_row_nodes
    rowptr, colind = _with_extra_blocks(rowptr0, colind0, nn, lambda k: (0, nn - 1, (7 * k + 3) % nn))
    assert rowptr[-1] > rowptr0[-1]
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, PAR, graph=(rowptr, colind))
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate()
    rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, orc.MAT_COUP_NEOHOOKE, PAR, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid,
                                         rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-12)
    assert ok and nbad == 0, (wK, wf)
    # the entries outside the block's own pattern are exactly zero
    own = set()
    for r in range(3 * nn):
        own |= {(r, c) for c in colind0[rowptr0[r]:rowptr0[r + 1]].tolist()}
    rows = np.repeat(np.arange(3 * nn), np.diff(rowptr))
    extra = np.array([(r, c) not in own for r, c in zip(rows.tolist(), colind.tolist())])
    assert extra.any() and not K[extra].any() and np.abs(K[~extra]).max() > 0
    asm.close()


def test_caller_graph_missing_an_entry_is_refused():
    """A caller graph that lacks a block the elements add into: the reference throws in assemble (a filled matrix cannot take
    new entries); here set_graph refuses the graph."""
    rm = b200.partition.box_partition(b200.mesh.HEX8, (3, 2, 2), 0, 1)
    rowptr0, colind0 = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    # drop the last node block of row node 0
    keep = np.ones(len(colind0), bool)
    for r in range(3):
        keep[rowptr0[r + 1] - 3:rowptr0[r + 1]] = False
    colind = colind0[keep]
    rowptr = rowptr0.copy()
    for r in range(3):
        rowptr[r + 1:] -= 3
    with pytest.raises(capi.B200SolidError):
        b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, PAR, graph=(rowptr, colind))
