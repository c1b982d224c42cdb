"""Row-element evaluation with export of the off-rank rows (FE_MATRIX mode of Core::LinAlg::SparseMatrix:
fe_assemble collects the rows a rank does not own, complete() ships them to their owners with an
Epetra_Export in Add mode, src/core/linalg/src/sparse/4C_linalg_sparsematrix.cpp:598-630,705-710).

The standard path evaluates every COLUMN element of a rank (all elements touching an owned node) and drops
the rows it does not own: the one-element ghost layer is evaluated twice.  Here a rank evaluates its ROW
elements only; contributions to rows of ghost nodes are accumulated in extra CSR rows behind the owned ones
("export rows", a mini graph of the ghost row node against the column nodes of the rank's row elements),
sent to the owner after the element kernels and added there.

This module is the host-side harness of that mode (the role Epetra's export plan plays in 4C):
  fe_view     row elements, export row nodes grouped by owner, extended row numbering
  fe_graph    extended CSR graph = owned rows (full pattern, from all column elements) + export rows
  fe_plan     per peer: the slice of the extended values / residual to send and, for what arrives, the CSR
              positions to add it to (b200s_set_row_export)
Partitions are deterministic functions of (rank, nranks), so a rank derives its peers' export rows by
evaluating their partition locally instead of exchanging index lists."""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class FeView:
    rm: object                   # partition.RankMesh of the rank
    ele_ids: np.ndarray          # indices of the row elements inside rm.conn
    conn: np.ndarray             # (n_row_eles, nn) int32 column-node lids
    export_nodes: np.ndarray     # column-node lids of the ghost nodes touched by row elements, ascending = grouped by owner
    export_owner: np.ndarray     # owner rank of every export node
    node_row_lid_ext: np.ndarray  # (n_col_nodes,) owned: row-node index, export node k: n_row_nodes + k, else -1

    @property
    def n_row_nodes_ext(self):
        return self.rm.n_row_nodes + len(self.export_nodes)


def fe_view(rm):
    assert rm.ele_is_row is not None
    ele_ids = np.nonzero(rm.ele_is_row)[0]
    conn = np.ascontiguousarray(rm.conn[ele_ids], np.int32)
    owner = np.full(rm.n_col_nodes, rm.rank, np.int64)
    for p, rcv in zip(rm.peers, rm.recv_col_nodes):
        owner[rcv] = p
    used = np.zeros(rm.n_col_nodes, bool)
    used[conn.ravel()] = True
    export_nodes = np.nonzero(used & (rm.node_row_lid < 0))[0].astype(np.int32)
    # column order of the ghosts is (owner, gid): ascending lids are grouped by owner
    assert np.all(np.diff(owner[export_nodes]) >= 0)
    ext = rm.node_row_lid.astype(np.int32).copy()
    ext[export_nodes] = rm.n_row_nodes + np.arange(len(export_nodes), dtype=np.int32)
    return FeView(rm, ele_ids, conn, export_nodes, owner[export_nodes], ext)


def node_graph(conn, row_of_node, n_row_nodes, n_col_nodes):
    """CSR graph of the dof rows of the nodes with row_of_node >= 0 against the column nodes of the elements that
    contain them (3 dofs per node, node-blocked, columns ascending): what the first assembly + FillComplete give
    (4C_linalg_sparsematrix.cpp:339-495).  Returns (rowptr int64 [3 n_row_nodes + 1], colind int32)."""
    conn = np.asarray(conn, np.int64)
    nn = conn.shape[1]
    r = np.asarray(row_of_node, np.int64)[conn]                      # (ne, nn)
    rows = np.repeat(r, nn, axis=1).ravel()
    cols = np.tile(conn, (1, nn)).ravel()
    m = rows >= 0
    key = np.unique(rows[m] * n_col_nodes + cols[m])
    rn, cn = key // n_col_nodes, key % n_col_nodes
    cnt = np.bincount(rn, minlength=n_row_nodes)
    node_start = np.concatenate(([0], np.cumsum(cnt)))
    rowlen = 3 * np.repeat(cnt, 3)                                    # entries per dof row
    rowptr = np.concatenate(([0], np.cumsum(rowlen))).astype(np.int64)
    colind = np.empty(rowptr[-1], np.int32)
    cdof = (3 * cn[:, None] + np.arange(3)[None, :]).astype(np.int32)  # (n_pairs, 3) column dofs of every pair
    for b in range(3):
        # dof row 3 k + b of node k holds the column dofs of its pairs in order
        start = rowptr[3 * rn + b] + 3 * (np.arange(len(rn)) - node_start[rn])
        colind[start[:, None] + np.arange(3)[None, :]] = cdof
    return rowptr, colind


def fe_graph(view):
    """Extended graph: owned rows with their full pattern (all column elements of the rank, as the reference's matrix
    has them) followed by the export rows (pattern of the row elements only)."""
    rm = view.rm
    rowptr_o, colind_o = node_graph(rm.conn, rm.node_row_lid, rm.n_row_nodes, rm.n_col_nodes)
    g = np.full(rm.n_col_nodes, -1, np.int64)
    g[view.export_nodes] = np.arange(len(view.export_nodes))
    rowptr_g, colind_g = node_graph(view.conn, g, len(view.export_nodes), rm.n_col_nodes)
    rowptr = np.concatenate((rowptr_o, rowptr_o[-1] + rowptr_g[1:])).astype(np.int64)
    colind = np.concatenate((colind_o, colind_g)).astype(np.int32)
    return rowptr, colind


@dataclass
class FePlan:
    n_rows_owned: int                     # dof rows the rank owns (the export rows follow them)
    nnz_owned: int
    peers: list = field(default_factory=list)
    send_val: list = field(default_factory=list)   # per peer (start, count) in the extended values
    send_f: list = field(default_factory=list)     # per peer (start, count) in the extended residual
    recv_pos: list = field(default_factory=list)   # per peer int64 positions in the owned values, in the sender's order
    recv_fpos: list = field(default_factory=list)  # per peer int32 positions in the owned residual


def _export_slices(view, rowptr):
    """owner -> ((value start, count), (f start, count)) of the rank's export rows."""
    out = {}
    n_row = view.rm.n_row_nodes
    for q in np.unique(view.export_owner):
        k = np.nonzero(view.export_owner == q)[0]
        a, b = int(k[0]), int(k[-1]) + 1
        assert np.array_equal(k, np.arange(a, b))
        v0, v1 = int(rowptr[3 * (n_row + a)]), int(rowptr[3 * (n_row + b)])
        out[int(q)] = ((v0, v1 - v0), (3 * (n_row + a), 3 * (b - a)), (a, b))
    return out


def fe_plan(mesh_of_rank, rank, nranks, graphs=None):
    """mesh_of_rank(r) -> RankMesh (deterministic partitioner).  Returns (view, rowptr, colind, plan) of `rank`."""
    graphs = {} if graphs is None else graphs

    def get(r):
        if r not in graphs:
            v = fe_view(mesh_of_rank(r))
            graphs[r] = (v,) + fe_graph(v)
        return graphs[r]

    view, rowptr, colind = get(rank)
    rm = view.rm
    plan = FePlan(3 * rm.n_row_nodes, int(rowptr[3 * rm.n_row_nodes]))
    mine = _export_slices(view, rowptr)
    gid_to_col = {int(g): k for k, g in enumerate(rm.node_gid)}
    # ranks that may hold export rows for me: those that ghost my nodes
    senders = [p for p, snd in zip(rm.peers, rm.send_row_nodes) if len(snd)]
    for p in sorted(set(list(mine) + senders)):
        rp, rf = np.zeros(0, np.int64), np.zeros(0, np.int32)
        if p in senders:
            pv, prowptr, pcolind = get(p)
            sl = _export_slices(pv, prowptr).get(rank)
            if sl is not None:
                (v0, nv), (f0, nf), (a, b) = sl
                pos = np.empty(nv, np.int64)
                fpos = np.empty(nf, np.int32)
                w = 0
                for k in range(a, b):
                    g = int(pv.rm.node_gid[pv.export_nodes[k]])
                    rl = int(rm.node_row_lid[gid_to_col[g]])
                    assert rl >= 0
                    for d in range(3):
                        prow = 3 * (pv.rm.n_row_nodes + k) + d
                        pc = pcolind[prowptr[prow]:prowptr[prow + 1]]
                        # sender's column dofs -> my column dofs (through the global node id)
                        pg = pv.rm.node_gid[pc // 3]
                        mc = 3 * np.fromiter((gid_to_col[int(x)] for x in pg), np.int64, len(pg)) + pc % 3
                        r0, r1 = rowptr[3 * rl + d], rowptr[3 * rl + d + 1]
                        idx = np.searchsorted(colind[r0:r1], mc)
                        assert np.array_equal(colind[r0:r1][idx], mc), "export row entry missing in the owner's graph"
                        pos[w:w + len(pc)] = r0 + idx
                        w += len(pc)
                        fpos[3 * (k - a) + d] = 3 * rl + d
                assert w == nv
                rp, rf = pos, fpos
        sv, sf = (mine[p][0], mine[p][1]) if p in mine else ((0, 0), (0, 0))
        if sv[1] == 0 and len(rp) == 0:
            continue
        plan.peers.append(int(p))
        plan.send_val.append(sv)
        plan.send_f.append(sf)
        plan.recv_pos.append(rp)
        plan.recv_fpos.append(rf)
    return view, rowptr, colind, plan
