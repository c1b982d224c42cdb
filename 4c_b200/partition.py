"""Rank-local view of a partitioned discretization, as 4C would hand it to the device path.

What 4C provides in a real run and this module reproduces for synthetic meshes:
  * node row map: a node shared by several ranks' row elements belongs to the LOWEST such rank
    (Core::Rebalance::build_graph, src/core/rebalance/src/4C_rebalance_graph_based.cpp:279-330),
    row lids ascend with the global id;
  * column elements: every element with at least one owned node; column nodes: their nodes
    (src/core/fem/src/discretization/4C_fem_discretization_partition.cpp:300-440);
  * column ordering (Epetra MakeColMap convention): owned ids first in row-map order,
    then ghosts grouped by owner rank, ascending global id inside a group -- used both for the
    discretization's column nodes and for the matrix-local column ids;
  * the ghost import of Discretization::set_state (4C_fem_discretization.cpp:653-661) as per-peer
    send / receive lists.
Structured boxes follow the grid generator's partition (src/core/io/src/4C_io_gridgenerator.cpp:
79-165) and are computed analytically per rank, so no rank ever builds the global mesh.
"""
from dataclasses import dataclass, field

import numpy as np

from . import mesh as _mesh


@dataclass
class RankMesh:
    celltype: int
    rank: int
    nranks: int
    X: np.ndarray              # (n_col_nodes,3) column nodes in column-map order
    conn: np.ndarray           # (n_col_eles,nn) int32 column-node lids
    node_gid: np.ndarray       # (n_col_nodes,) global node ids
    node_row_lid: np.ndarray   # (n_col_nodes,) int32 row-node index or -1 (ghost)
    node_col_lid: np.ndarray   # (n_col_nodes,) int32 matrix-local column-node index (= arange)
    n_row_nodes: int
    n_row_eles: int            # elements this rank "owns" for the throughput metric
    ele_gid: np.ndarray        # (n_col_eles,)
    peers: list = field(default_factory=list)       # neighbour ranks
    send_row_nodes: list = field(default_factory=list)  # per peer: row-node indices to send
    recv_col_nodes: list = field(default_factory=list)  # per peer: column-node lids to fill
    ele_is_row: np.ndarray = None  # (n_col_eles,) uint8: row (owned) elements; None = all

    @property
    def n_col_nodes(self):
        return len(self.X)

    @property
    def n_col_eles(self):
        return len(self.conn)


def box_subdivisions(n, nranks):
    """Subdivision of the element box into nranks sub-boxes (4C_io_gridgenerator.cpp:91-140)."""
    factors, p, fac = [], nranks, 2
    while fac < p + 1:
        if p % fac == 0:
            factors.append(fac)
            p //= fac
        else:
            fac += 1
    sub = [1, 1, 1]
    for f in reversed(factors):
        ratios = [n[d] / sub[d] for d in range(3)]
        if ratios[0] >= ratios[1] and ratios[0] >= ratios[2]:
            sub[0] *= f
        elif ratios[1] >= ratios[0] and ratios[1] >= ratios[2]:
            sub[1] *= f
        else:
            sub[2] *= f
    ranges = [[max(0, min(n[d], int(round(i * n[d] / sub[d])))) for i in range(sub[d] + 1)]
              for d in range(3)]
    return sub, ranges


def _section(rank, sub):
    return (rank % sub[0], (rank // sub[0]) % sub[1], rank // (sub[0] * sub[1]))


def _axis_sets(celltype, nd, rng_d, s):
    """Per axis for section s: owned node index range, column node range, column element range."""
    step = 1 if celltype == _mesh.HEX8 else 2
    e0, e1 = rng_d[s], rng_d[s + 1]
    own_lo = step * e0 + (1 if e0 > 0 else 0)      # lower interface plane belongs to the lower rank
    own_hi = step * e1                              # inclusive
    ce1 = min(e1 + 1, nd)                           # one ghost element layer on the upper side
    col_lo, col_hi = step * e0, step * ce1          # inclusive
    return (own_lo, own_hi), (col_lo, col_hi), (e0, ce1)


def box_partition(celltype, n, rank, nranks, box=(1.0, 1.0, 1.0)):
    """Rank-local structured hex8/hex27 mesh for the grid generator's box partition."""
    sub, ranges = box_subdivisions(n, nranks)
    if any(ranges[d][k + 1] <= ranges[d][k] for d in range(3) for k in range(sub[d])):
        # a section without elements has no interface plane of its own: the ownership rule below ("lower interface plane
        # belongs to the lower rank") is not defined for it
        raise ValueError(f"box_partition: {n} elements cannot be cut into {sub} non-empty sections ({nranks} ranks)")
    sec = _section(rank, sub)
    (gx, gy, gz), step = _mesh.hex_grid_dims(celltype, n)
    own, col, cel = zip(*[_axis_sets(celltype, n[d], ranges[d], sec[d]) for d in range(3)])

    # column elements, ascending global id
    ex = np.arange(cel[0][0], cel[0][1])
    ey = np.arange(cel[1][0], cel[1][1])
    ez = np.arange(cel[2][0], cel[2][1])
    EZ, EY, EX = np.meshgrid(ez, ey, ex, indexing="ij")
    EX, EY, EZ = EX.ravel(), EY.ravel(), EZ.ravel()
    ele_gid = (EZ * n[1] + EY) * n[0] + EX
    conn_g = _mesh.hex_element_nodes(celltype, n, EX, EY, EZ)
    n_row_eles = int(np.prod([ranges[d][sec[d] + 1] - ranges[d][sec[d]] for d in range(3)]))

    # column nodes = nodes of the column elements (hex8: the whole index box; hex27: as well)
    ix = np.arange(col[0][0], col[0][1] + 1)
    iy = np.arange(col[1][0], col[1][1] + 1)
    iz = np.arange(col[2][0], col[2][1] + 1)
    IZ, IY, IX = np.meshgrid(iz, iy, ix, indexing="ij")
    IX, IY, IZ = IX.ravel(), IY.ravel(), IZ.ravel()
    gid = (IZ * gy + IY) * gx + IX

    def owner_section(i, d):
        # node index i along axis d -> section of the lowest element touching it
        e = np.maximum((i - 1) // step, 0) if step == 1 else np.maximum((i - 1) // 2, 0)
        e = np.minimum(e, n[d] - 1)
        return np.searchsorted(np.asarray(ranges[d][1:]), e, side="right")

    osx, osy, osz = owner_section(IX, 0), owner_section(IY, 1), owner_section(IZ, 2)
    owner = osx + sub[0] * (osy + sub[1] * osz)
    mine = owner == rank
    # column-map order: owned ascending gid, then ghosts by (owner, gid)
    order = np.lexsort((gid, np.where(mine, -1, owner)))
    gid, owner, mine = gid[order], owner[order], mine[order]
    n_row = int(mine.sum())
    node_row_lid = np.where(mine, np.arange(len(gid)), -1).astype(np.int32)
    # global id -> column lid (sorted lookup)
    srt = np.argsort(gid, kind="stable")
    conn = srt[np.searchsorted(gid[srt], conn_g)].astype(np.int32)
    X = _mesh.hex_node_coords(celltype, n, gid, box)

    rm = RankMesh(celltype, rank, nranks, X, conn, gid, node_row_lid,
                  np.arange(len(gid), dtype=np.int32), n_row, n_row_eles, ele_gid)
    # row elements = the rank's own section of the element grid
    lo = [ranges[d][sec[d]] for d in range(3)]
    hi = [ranges[d][sec[d] + 1] for d in range(3)]
    rm.ele_is_row = ((EX >= lo[0]) & (EX < hi[0]) & (EY >= lo[1]) & (EY < hi[1]) & (EZ >= lo[2]) &
                     (EZ < hi[2])).astype(np.uint8)
    assert int(rm.ele_is_row.sum()) == n_row_eles

    # halo lists: my ghosts grouped by owner (receive), and for every neighbour q the nodes I own
    # that lie in q's column-node box (send), both ascending in gid = the order q stores them
    ghost_owner = owner[~mine]
    for p in np.unique(ghost_owner):
        rm.peers.append(int(p))
    nbr = set(rm.peers)
    own_box = [np.arange(own[d][0], own[d][1] + 1) for d in range(3)]
    for q in range(nranks):
        if q == rank:
            continue
        qsec = _section(q, sub)
        if any(abs(qsec[d] - sec[d]) > 1 for d in range(3)):
            continue
        qown, qcol, _ = zip(*[_axis_sets(celltype, n[d], ranges[d], qsec[d]) for d in range(3)])
        inter = [own_box[d][(own_box[d] >= qcol[d][0]) & (own_box[d] <= qcol[d][1])]
                 for d in range(3)]
        if any(len(v) == 0 for v in inter):
            continue
        nbr.add(q)
    rm.peers = sorted(nbr)
    gid_to_row = dict()  # only for the (small) interface sets
    row_gid = gid[:n_row]
    for q in rm.peers:
        qsec = _section(q, sub)
        qown, qcol, _ = zip(*[_axis_sets(celltype, n[d], ranges[d], qsec[d]) for d in range(3)])
        inter = [own_box[d][(own_box[d] >= qcol[d][0]) & (own_box[d] <= qcol[d][1])]
                 for d in range(3)]
        if all(len(v) for v in inter):
            JZ, JY, JX = np.meshgrid(inter[2], inter[1], inter[0], indexing="ij")
            sg = np.sort(((JZ * gy + JY) * gx + JX).ravel())
            rm.send_row_nodes.append(np.searchsorted(row_gid, sg).astype(np.int32))
        else:
            rm.send_row_nodes.append(np.zeros(0, np.int32))
        rm.recv_col_nodes.append((n_row + np.nonzero(ghost_owner == q)[0]).astype(np.int32))
    return rm


def generic_partition(celltype, X, conn, node_owner, rank, nranks, ele_owner=None):
    """Same rank-local view for an arbitrary mesh and node-ownership array (tests, tet10)."""
    conn = np.asarray(conn, np.int64)
    node_owner = np.asarray(node_owner)
    touch = (node_owner[conn] == rank).any(axis=1)
    ele_gid = np.nonzero(touch)[0]
    cg = conn[ele_gid]
    gid = np.unique(cg)
    owner = node_owner[gid]
    mine = owner == rank
    order = np.lexsort((gid, np.where(mine, -1, owner)))
    gid, owner, mine = gid[order], owner[order], mine[order]
    n_row = int(mine.sum())
    srt = np.argsort(gid, kind="stable")
    lconn = srt[np.searchsorted(gid[srt], cg)].astype(np.int32)
    node_row_lid = np.where(mine, np.arange(len(gid)), -1).astype(np.int32)
    if ele_owner is None:
        # an element is counted for the rank owning most of its nodes (ties: lowest rank)
        own_cnt = np.stack([(node_owner[conn] == r).sum(axis=1) for r in range(nranks)], axis=1)
        ele_owner = own_cnt.argmax(axis=1)
    rm = RankMesh(celltype, rank, nranks, np.asarray(X, np.float64)[gid], lconn, gid, node_row_lid,
                  np.arange(len(gid), dtype=np.int32), n_row, int((ele_owner == rank).sum()), ele_gid)
    rm.ele_is_row = (np.asarray(ele_owner)[ele_gid] == rank).astype(np.uint8)
    ghost_owner = owner[~mine]
    row_gid = gid[:n_row]
    peers = set(int(p) for p in np.unique(ghost_owner))
    # who ghosts my nodes: ranks q != me that own a node of an element containing one of my nodes
    sends = {}
    for q in range(nranks):
        if q == rank:
            continue
        tq = (node_owner[conn] == q).any(axis=1)
        gq = np.unique(conn[tq])
        sg = gq[node_owner[gq] == rank]
        if len(sg):
            sends[q] = np.searchsorted(row_gid, sg).astype(np.int32)
            peers.add(q)
    rm.peers = sorted(peers)
    for q in rm.peers:
        rm.send_row_nodes.append(sends.get(q, np.zeros(0, np.int32)))
        rm.recv_col_nodes.append((n_row + np.nonzero(ghost_owner == q)[0]).astype(np.int32))
    return rm


def serial_mesh(celltype, X, conn):
    """Single-rank view: every node owned, column order = global id order."""
    n = len(X)
    ids = np.arange(n, dtype=np.int32)
    return RankMesh(celltype, 0, 1, np.asarray(X, np.float64), np.asarray(conn, np.int32),
                    np.arange(n, dtype=np.int64), ids.copy(), ids.copy(), n, len(conn),
                    np.arange(len(conn), dtype=np.int64))
