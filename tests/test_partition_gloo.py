"""Multi-rank host logic on CPU (gloo, world_size 2 and 4): the rank-local discretization views of
4c_b200.partition (row/column maps, ghost import lists) drive a real point-to-point halo exchange and
an owner-computes assembly whose rows, put together, equal the serial assembly.  The assembly itself
is done by the oracle here (no GPU); the same lists feed b200s_set_halo on the device path
(tests/test_gpu_multirank.py)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def halo_exchange(rm, u_row):
    """Discretization::set_state ghost import (4C_fem_discretization.cpp:653-661) with the lists of
    partition.RankMesh over torch.distributed point-to-point: returns u on the column map."""
    u_col = np.zeros((rm.n_col_nodes, 3))
    u_col[:rm.n_row_nodes] = u_row.reshape(-1, 3)
    reqs, recv_bufs = [], []
    for p, snd, rcv in zip(rm.peers, rm.send_row_nodes, rm.recv_col_nodes):
        if len(snd):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(u_col[snd]).ravel().copy()), dst=p))
        if len(rcv):
            buf = torch.empty(3 * len(rcv), dtype=torch.float64)
            recv_bufs.append((rcv, buf))
            reqs.append(dist.irecv(buf, src=p))
    for r in reqs:
        r.wait()
    for rcv, buf in recv_bufs:
        u_col[rcv] = buf.numpy().reshape(-1, 3)
    return u_col


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b200 = importlib.import_module("4c_b200")
    from oracle import oracle as orc
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, rank, world)
    u_exact = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / max(n))
    u_col = halo_exchange(rm, u_exact[:rm.n_row_nodes].ravel())
    ok_halo = np.array_equal(u_col, u_exact)
    rowptr, colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    K, f, nbad = orc.evaluate_assemble(orc.HEX8, orc.MAT_COUP_NEOHOOKE, [200.0, 0.25], rm.conn, rm.X.ravel(),
                                       u_col.ravel(), rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes,
                                       rowptr, colind)
    # global (row gid, col gid) -> value triplets of this rank's rows
    rows = np.repeat(np.arange(3 * rm.n_row_nodes), np.diff(rowptr))
    g_row = 3 * rm.node_gid[rows // 3] + rows % 3
    g_col = 3 * rm.node_gid[colind // 3] + colind % 3
    np.savez(os.path.join(out, f"rank{rank}.npz"), ok_halo=ok_halo, g_row=g_row, g_col=g_col, K=K,
             f=f, f_gid=(3 * rm.node_gid[:rm.n_row_nodes, None] + np.arange(3)).ravel(), nbad=nbad,
             n_row_eles=rm.n_row_eles, n_col_eles=rm.n_col_eles)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, (6, 4, 3)), (4, (6, 5, 3)), (2, (2, 1, 1))])
def test_owner_computes_rows_over_gloo_equal_serial(world, n, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    b200 = importlib.import_module("4c_b200")
    from oracle import oracle as orc
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / max(n))
    rowptr, colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    K, f, _ = orc.evaluate_assemble(orc.HEX8, orc.MAT_COUP_NEOHOOKE, [200.0, 0.25], rm.conn, rm.X.ravel(),
                                    u.ravel(), rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    ndof = 3 * rm.n_row_nodes
    rows = np.repeat(np.arange(ndof), np.diff(rowptr))
    serial = dict(zip((rows.astype(np.int64) * ndof + colind).tolist(), K.tolist()))
    seen_rows, n_entries, row_eles = set(), 0, 0
    f_all = np.full(ndof, np.nan)
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert bool(d["ok_halo"]) and int(d["nbad"]) == 0
        keys = d["g_row"].astype(np.int64) * ndof + d["g_col"]
        assert len(np.unique(keys)) == len(keys)
        ref = np.array([serial[k] for k in keys.tolist()])   # KeyError = entry outside the serial pattern
        scale = np.abs(K).max()
        assert np.abs(d["K"] - ref).max() <= 1e-12 * scale
        assert not (set(d["f_gid"].tolist()) & seen_rows)    # every row owned exactly once
        seen_rows |= set(d["f_gid"].tolist())
        f_all[d["f_gid"]] = d["f"]
        n_entries += len(keys)
        row_eles += int(d["n_row_eles"])
    assert len(seen_rows) == ndof and n_entries == len(K)
    assert np.abs(f_all - f).max() <= 1e-12 * np.abs(f).max()
    assert row_eles == n[0] * n[1] * n[2]


def test_box_partition_refuses_empty_sections():
    """More ranks than elements along the cut axis: refused loudly (the halo lists of an empty section are undefined)."""
    b200 = importlib.import_module("4c_b200")
    for n, world in (((1, 1, 1), 2), ((2, 1, 1), 4)):
        with pytest.raises(ValueError):
            b200.partition.box_partition(b200.mesh.HEX8, n, 0, world)
    rm = b200.partition.box_partition(b200.mesh.HEX8, (2, 1, 1), 1, 2)
    assert rm.n_row_eles == 1 and rm.n_col_eles == 1
