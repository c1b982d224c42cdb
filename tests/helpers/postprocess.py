"""TEST INFRASTRUCTURE: Gauss point -> node extrapolation of element stresses as the reference does it
for its nodal stress results (src/core/fem/src/general/utils/4C_fem_general_utils_gauss_point_extrapolation.cpp):
  base cell by number of Gauss points (:20-37: hex 8 GP -> hex8, 27 GP -> hex27, tet 4 GP -> tet4),
  projection = inverse of the base shape functions evaluated at the Gauss points (:150-176),
  extension to the element's nodes with the base shape functions at the node coordinates (:207-258),
  assembly with nodal averaging 1 / (number of adjacent elements) (:260-277).
Used to compare the stress path with the RESULT DESCRIPTION known answers of the reference's inputs."""
import numpy as np

from oracle import oracle as orc


def _tet4(xi):
    r, s, t = xi
    return np.array([1.0 - r - s - t, r, s, t])


def extrapolation_matrix(ct):
    gxi, _ = orc.gauss(ct)
    nodes = orc.node_param_coords(ct)
    if ct == orc.TET10:
        base = _tet4
    else:
        base = lambda xi: orc.shape(ct, xi)[0]  # hex8 with 8 GP / hex27 with 27 GP: base == cell
    at_gp = np.array([base(x) for x in gxi])         # (ngp, nbase)
    assert at_gp.shape[0] == at_gp.shape[1]
    gp_to_base = np.linalg.inv(at_gp)                # (nbase, ngp)
    base_to_nodes = np.array([base(x) for x in nodes])  # (nn, nbase)
    return base_to_nodes @ gp_to_base                # (nn, ngp)


def nodal_stresses(ct, conn, gp_stress, n_nodes):
    """gp_stress [n_ele, ngp, 6] -> nodal [n_nodes, 6] (averaged over adjacent elements)."""
    A = extrapolation_matrix(ct)
    out = np.zeros((n_nodes, 6))
    cnt = np.zeros(n_nodes)
    for e, ce in enumerate(np.asarray(conn)):
        cnt[ce] += 1
    for e, ce in enumerate(np.asarray(conn)):
        out[ce] += (A @ gp_stress[e]) / cnt[ce][:, None]
    return out
