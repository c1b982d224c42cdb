"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Row-sampled parity check of a device-resident assembly against the CPU oracle, for sizes at which the
whole matrix neither fits the host comfortably nor can be evaluated by the oracle in seconds (BASELINE
configs 3, 4 and 5: nnz per rank > 2^31).

The rows of a set of owned nodes are complete once every column element touching one of those nodes has
been evaluated -- exactly the owner-computes rule of SparseMatrix::assemble
(4C_linalg_sparsematrix.cpp:386: rows with lmrowowner != myrank are skipped).  So the oracle evaluates
only the elements touching the sampled nodes, assembles the sampled rows in the SAME column numbering,
and the result is compared entry by entry with the rows fetched through b200s_get_value_rows /
b200s_get_f_rows.  Used by tests/ and by bench.py's `parity` record (checker only, after the timed
region).
"""
import numpy as np

from . import oracle as orc


def pick_nodes(rm, n_random=150, n_interface=50, seed=0):
    """Owned nodes to sample: random ones plus nodes of column elements that also hold ghost nodes (their
    rows receive contributions computed from halo-imported displacements)."""
    rng = np.random.default_rng(seed)
    owned = np.nonzero(rm.node_row_lid >= 0)[0]
    pick = rng.choice(owned, min(n_random, len(owned)), replace=False)
    ghost = rm.node_row_lid < 0
    if ghost.any() and n_interface > 0:
        touch = ghost[rm.conn].any(axis=1)
        iface = np.unique(rm.conn[touch].ravel())
        iface = iface[rm.node_row_lid[iface] >= 0]
        if len(iface):
            pick = np.concatenate([pick, rng.choice(iface, min(n_interface, len(iface)), replace=False)])
    return np.unique(pick)


def check(ctx, rm, celltype, matkind, params, u_col, nodes, rtol=1e-12, mue_per_ele=None, tech=orc.TECH_NONE,
          nthreads=4, floor=1e-6):
    """ctx: capi.Context after an evaluate with K resident.  u_col: (n_col_nodes,3) displacement incl. ghosts.
    nodes: column-node lids of OWNED nodes.  Returns a dict:
      ok, worst_blockrow (|a-b| / (rtol*max-abs of the node block row), the SURVEY 8(d) norm), worst_entry_rel
      (pure |a-b|/max(|a|,|b|) over entries above floor*block-row scale), worst_f, pattern_ok, rows, entries."""
    nodes = np.sort(np.asarray(nodes))
    keep = np.zeros(rm.n_col_nodes, bool)
    keep[nodes] = True
    sel = keep[rm.conn].any(axis=1)
    sub_row = np.where(keep, np.cumsum(keep) - 1, -1).astype(np.int32)
    nsub = len(nodes)
    conn_s = np.ascontiguousarray(rm.conn[sel])
    o_rowptr, o_colind = orc.build_graph(celltype, conn_s, sub_row, rm.node_col_lid, nsub)
    mue = None if mue_per_ele is None else np.ascontiguousarray(np.asarray(mue_per_ele)[sel])
    rK, rf, nbad = orc.evaluate_assemble(celltype, matkind, params, conn_s, np.ascontiguousarray(rm.X).ravel(),
                                         np.ascontiguousarray(u_col).ravel(), sub_row, rm.node_col_lid, nsub,
                                         o_rowptr, o_colind, mue_per_ele=mue, nthreads=nthreads, tech=tech)
    worst_blk = worst_rel = worst_f = 0.0
    pattern_ok = True
    entries = 0
    fscale = None
    dev_f = np.zeros(3 * nsub)
    for k, node in enumerate(nodes):
        r0 = 3 * int(rm.node_row_lid[node])
        rp, ci, v = ctx.get_value_rows(r0, r0 + 3)
        dev_f[3 * k:3 * k + 3] = ctx.get_f_rows(r0, r0 + 3)
        lo, hi = o_rowptr[3 * k], o_rowptr[3 * k + 3]
        if hi - lo != len(v) or not np.array_equal(ci, o_colind[lo:hi]) or \
                not np.array_equal(rp, o_rowptr[3 * k:3 * k + 4] - lo):
            pattern_ok = False
            continue
        ref = rK[lo:hi]
        scale = max(np.abs(v).max(), np.abs(ref).max(), 1e-300)
        d = np.abs(v - ref)
        worst_blk = max(worst_blk, float(d.max() / (rtol * scale)))
        big = np.maximum(np.abs(v), np.abs(ref)) > floor * scale
        if big.any():
            worst_rel = max(worst_rel, float((d[big] / np.maximum(np.abs(v[big]), np.abs(ref[big]))).max()))
        entries += len(v)
    fscale = max(np.abs(dev_f).max(), np.abs(rf).max(), 1e-300)
    worst_f = float(np.abs(dev_f - rf).max() / (rtol * fscale))
    finite = bool(np.isfinite(dev_f).all())
    ok = bool(pattern_ok and finite and nbad == 0 and worst_blk <= 1.0 and worst_f <= 1.0)
    return {"ok": ok, "pattern_ok": bool(pattern_ok), "worst_blockrow_ratio": worst_blk,
            "worst_entry_rel": worst_rel, "worst_f_ratio": worst_f, "rtol": rtol, "entry_floor": floor,
            "nodes": int(nsub), "entries": int(entries), "elements_evaluated_by_oracle": int(sel.sum())}
