"""Small serial-mesh helpers for the tests (single rank: every node is owned; matrix-local column
ids equal node ids)."""
import numpy as np


def serial_maps(n_nodes):
    ids = np.arange(n_nodes, dtype=np.int32)
    return ids.copy(), ids.copy()  # node_row_lid, node_col_lid


def csr_to_dense(rowptr, colind, values, ncols):
    n = len(rowptr) - 1
    K = np.zeros((n, ncols))
    for r in range(n):
        K[r, colind[rowptr[r]:rowptr[r + 1]]] += values[rowptr[r]:rowptr[r + 1]]
    return K
