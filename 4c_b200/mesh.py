"""Synthetic meshes and displacement states for the hot-path workloads (BASELINE.json configs).

Structured hex meshes follow the numbering of the reference's grid generator
(src/core/io/src/4C_io_gridgenerator.cpp:375-425: element id (ez*ny+ey)*nx+ex, local node k sits at
half-grid offset node_ref[k]+1); hex8 uses contiguous node ids on the (n+1)^3 corner grid
(SURVEY.md 8d config 2), hex27 the full (2n+1)^3 grid.  The tet10 mesh is unstructured: every cell
of a hex grid is split into 6 tets, nodes are jittered and nodes / elements are renumbered with a
seeded permutation (config 4).
"""
import numpy as np

HEX8, HEX27, TET10, TET4, HEX20, WEDGE6, PYRAMID5, HEX18 = 0, 1, 2, 3, 4, 5, 6, 7
NODES = {HEX8: 8, HEX27: 27, TET10: 10, TET4: 4, HEX20: 20, WEDGE6: 6, PYRAMID5: 5, HEX18: 18}

# node order of hex27 in parameter space (first 8 = hex8);
# 4C_fem_general_utils_local_connectivity_matrices.hpp:291
HEX_REF = np.array(
    [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1],
     [-1, 1, 1], [0, -1, -1], [1, 0, -1], [0, 1, -1], [-1, 0, -1], [-1, -1, 0], [1, -1, 0],
     [1, 1, 0], [-1, 1, 0], [0, -1, 1], [1, 0, 1], [0, 1, 1], [-1, 0, 1], [0, 0, -1], [0, -1, 0],
     [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, 0, 1], [0, 0, 0]], dtype=np.int64)


def hex_grid_dims(celltype, n):
    """Nodes per direction of the node grid and half-grid step of an element."""
    nx, ny, nz = n
    if celltype == HEX8:
        return (nx + 1, ny + 1, nz + 1), 1
    return (2 * nx + 1, 2 * ny + 1, 2 * nz + 1), 2


def hex_element_nodes(celltype, n, ex, ey, ez):
    """Global node ids (len(e), nn) of elements at integer positions ex,ey,ez (arrays)."""
    (gx, gy, gz), step = hex_grid_dims(celltype, n)
    nn = NODES[celltype]
    off = (HEX_REF[:nn] + 1) // (2 // step)  # hex8: 0/1, hex27: 0/1/2
    ix = step * np.asarray(ex, np.int64)[:, None] + off[None, :, 0]
    iy = step * np.asarray(ey, np.int64)[:, None] + off[None, :, 1]
    iz = step * np.asarray(ez, np.int64)[:, None] + off[None, :, 2]
    return (iz * gy + iy) * gx + ix


def hex_node_coords(celltype, n, gid, box=(1.0, 1.0, 1.0)):
    (gx, gy, gz), _ = hex_grid_dims(celltype, n)
    gid = np.asarray(gid, np.int64)
    i, j, k = gid % gx, (gid // gx) % gy, gid // (gx * gy)
    return np.stack([i * (box[0] / (gx - 1)), j * (box[1] / (gy - 1)), k * (box[2] / (gz - 1))],
                    axis=1)


def structured_hex(celltype, n, box=(1.0, 1.0, 1.0)):
    """Whole mesh on one rank: returns X (n_nodes,3), conn (n_ele,nn) int32."""
    nx, ny, nz = n
    e = np.arange(nx * ny * nz, dtype=np.int64)
    conn = hex_element_nodes(celltype, n, e % nx, (e // nx) % ny, e // (nx * ny))
    (gx, gy, gz), _ = hex_grid_dims(celltype, n)
    X = hex_node_coords(celltype, n, np.arange(gx * gy * gz), box)
    return X, conn.astype(np.int32)


# tet10: edge (a,b) of the 4 corners -> mid node slot; node order
# 4C_fem_general_utils_local_connectivity_matrices.hpp:314, shape functions ...shapefunctions.hpp:252-267
TET10_EDGES = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]
# 6 tets of a hex sharing the main diagonal 0-6 (positive orientation)
HEX_TO_TETS = [(0, 1, 2, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6), (0, 4, 5, 6), (0, 5, 1, 6)]


def unstructured_tet10(n, seed=4321, jitter=0.2, renumber=True, box=(1.0, 1.0, 1.0)):
    """n^3 hex cells split into 6 tet10 each; straight-edge mid nodes; seeded jitter/renumbering."""
    rng = np.random.default_rng(seed)
    Xc, hexes = structured_hex(HEX8, (n, n, n), box)
    h = box[0] / n
    interior = np.all((Xc > 1e-12) & (Xc < np.array(box) - 1e-12), axis=1)
    Xc = Xc + interior[:, None] * (jitter * h * rng.uniform(-1, 1, Xc.shape)) * 0.5
    tets = np.concatenate([hexes[:, list(t)] for t in HEX_TO_TETS], axis=0).astype(np.int64)
    # positive volume check / fix orientation
    a, b, c, d = (Xc[tets[:, q]] for q in range(4))
    vol = np.einsum("ij,ij->i", np.cross(b - a, c - a), d - a)
    flip = vol < 0
    tets[flip, 1], tets[flip, 2] = tets[flip, 2].copy(), tets[flip, 1].copy()
    nc = len(Xc)
    # unique edges -> mid nodes
    ed = np.concatenate([np.sort(tets[:, list(e)], axis=1) for e in TET10_EDGES], axis=0)
    key = ed[:, 0] * nc + ed[:, 1]
    uk, inv = np.unique(key, return_inverse=True)
    mid = 0.5 * (Xc[uk // nc] + Xc[uk % nc])
    X = np.concatenate([Xc, mid], axis=0)
    ne = len(tets)
    conn = np.concatenate([tets, (nc + inv).reshape(6, ne).T], axis=1)
    if renumber:
        pn = rng.permutation(len(X))  # new id of old node
        Xn = np.empty_like(X)
        Xn[pn] = X
        X = Xn
        conn = pn[conn]
        conn = conn[rng.permutation(ne)]
    return X, conn.astype(np.int32)


def unstructured_tet4(n, seed=4321, jitter=0.2, renumber=True, box=(1.0, 1.0, 1.0)):
    """The corner-node (linear) version of unstructured_tet10: same cells, jitter and renumbering recipe."""
    X10, c10 = unstructured_tet10(n, seed, jitter, renumber=False, box=box)
    nc = (n + 1) ** 3                      # corner nodes come first in the un-renumbered tet10 mesh
    X, conn = X10[:nc], c10[:, :4].astype(np.int64)
    if renumber:
        rng = np.random.default_rng(seed + 1)
        pn = rng.permutation(nc)
        Xn = np.empty_like(X)
        Xn[pn] = X
        X, conn = Xn, pn[conn][rng.permutation(len(conn))]
    return X, conn.astype(np.int32)


def structured_hex20(n, box=(1.0, 1.0, 1.0)):
    """Serendipity version of the structured hex27 mesh: the first 20 nodes of every hex27 element (corners and
    mid-edge nodes, 4C_fem_general_utils_local_connectivity_matrices.hpp), unused grid nodes removed."""
    X27, c27 = structured_hex(HEX27, n, box)
    c = c27[:, :20]
    used = np.unique(c)
    new = np.full(len(X27), -1, np.int64)
    new[used] = np.arange(len(used))
    return X27[used], new[c].astype(np.int32)


def structured_hex18(n, box=(1.0, 1.0, 1.0)):
    """hex18 version of the structured hex27 mesh: the bottom and top quad9 of every hex27 element (corners, mid-sides,
    face centre; 4C_fem_general_utils_local_connectivity_matrices.hpp:174-179), unused grid nodes removed."""
    X27, c27 = structured_hex(HEX27, n, box)
    c = c27[:, [0, 1, 2, 3, 8, 9, 10, 11, 20, 4, 5, 6, 7, 16, 17, 18, 19, 25]]
    used = np.unique(c)
    new = np.full(len(X27), -1, np.int64)
    new[used] = np.arange(len(used))
    return X27[used], new[c].astype(np.int32)


def structured_wedge6(n, box=(1.0, 1.0, 1.0), jitter=0.15, seed=77):
    """Every cell of the structured hex8 grid split into two wedges along the diagonal of its bottom face (node order of
    the wedge6 shape functions: bottom triangle, then the top triangle above it), interior nodes jittered."""
    X, c8 = structured_hex(HEX8, n, box)
    X = _jitter_interior(X, n, box, jitter, seed)
    w = np.concatenate([c8[:, [0, 1, 2, 4, 5, 6]], c8[:, [0, 2, 3, 4, 6, 7]]])
    return X, np.ascontiguousarray(w, np.int32)


def structured_pyramid5(n, box=(1.0, 1.0, 1.0), jitter=0.15, seed=78):
    """Every cell of the structured hex8 grid split into six pyramids with the cell centre as the apex; the base
    quadrilaterals are ordered so that the mapping has a positive Jacobian (outward faces of the hex8 seen from the apex)."""
    X, c8 = structured_hex(HEX8, n, box)
    X = _jitter_interior(X, n, box, jitter, seed)
    cen = X[c8].mean(axis=1)
    apex = len(X) + np.arange(len(c8))
    faces = [[0, 1, 2, 3], [4, 7, 6, 5], [0, 4, 5, 1], [1, 5, 6, 2], [2, 6, 7, 3], [3, 7, 4, 0]]
    p = np.concatenate([np.column_stack([c8[:, f], apex]) for f in faces])
    return np.concatenate([X, cen]), np.ascontiguousarray(p, np.int32)


def _jitter_interior(X, n, box, jitter, seed):
    n = (n,) * 3 if np.isscalar(n) else tuple(n)
    h = np.array(box, float) / np.array(n, float)
    rng = np.random.default_rng(seed)
    X = X.copy()
    inner = np.all((X > 1e-12) & (X < np.array(box) - 1e-12), axis=1)
    X[inner] += jitter * h * rng.uniform(-1, 1, (int(inner.sum()), 3))
    return X


def _hash_uniform(idx, seed):
    """Deterministic U(-1,1) from integer ids (splitmix64), identical on every rank."""
    z = (np.asarray(idx, np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def displacement_field(X, gid, L=1.0, h=0.01, amp=0.02, noise=1e-3, seed=1234):
    """Smooth finite-strain field + per-node noise: det F stays in (0.8, 1.2) and no
    two elements are identical.  gid: global node ids (noise is a function of the id only)."""
    X = np.asarray(X, np.float64)
    x, y, z = X[:, 0] / L, X[:, 1] / L, X[:, 2] / L
    u = amp * L * np.stack([np.sin(2 * np.pi * y) * z, np.sin(2 * np.pi * z) * x,
                            np.sin(2 * np.pi * x) * y], axis=1)
    gid = np.asarray(gid, np.uint64)
    with np.errstate(over="ignore"):
        nz = np.stack([_hash_uniform(gid * np.uint64(3) + np.uint64(d), seed) for d in range(3)],
                      axis=1)
    return u + noise * h * nz
