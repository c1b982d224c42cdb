#!/usr/bin/env python
"""bench.py -- solid element evaluate+assemble throughput (FP64) on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one struct_calc_nlnstiff evaluate of the hot path over the rank's element block
(BASELINE.json config 2 at N=1: synthetic structured hex8 Neo-Hooke cube, 100^3 = 1M elements):
ghost import (N>1), zero K/f, element kernel with scatter-add into the CSR values and residual.

  value : whole-job elements/s (row elements of all ranks / max-over-ranks device time); inputs
          (u on the dof row map) resident in HBM, outputs (K values, f) left in HBM.
  e2e   : the same metric through the blocking C-ABI call b200s_evaluate with HOST (pinned)
          buffers -- u copied host->device, K values and f copied device->host every step.
  roofline / cpu_baseline : see DESIGN.md "Measurement".

--impl reference times the reference's own CPU arithmetic (oracle/_ref, built on the reference's
header-only tensor/shape-function code; falls back to the oracle port) on all host threads.
Nothing here reads /root/reference.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "elements/s"
L2_BYTES = 126e6
# Workloads = BASELINE.json configs 2 (default, the one the metric is quoted on), 3 and 4.  flops / bytes per
# element: SURVEY.md 8(d) / Appendix B algorithmic model (62 n^2 + 228 n + 1183 + M per Gauss point).
WORKLOADS = {
    "hex8": dict(config="2", cell="HEX8", mat="MAT_COUP_NEOHOOKE", par=[200.0, 0.25], matname="CoupNeoHooke (E=200, nu=0.25)",
                 flops=63.8e3, bytes=2304.0, kernel="k_hex8_eval_assemble<COUP_NEOHOOKE,STIFF>", default_n=100,
                 metric="hex8 Neo-Hooke element evaluate+assemble (force + tangent stiffness into CSR), FP64"),
    # Section 8(f) row 1: F-bar hex8 on the config-2 mesh.  Model = hex8 model + the F-bar terms of the
    # reference per Gauss point: rank-1 update of the 24x24 matrix (2*24^2), Bop^T ccg / Bop^T S (2*2*6*24),
    # H-operator (2*2*3*24), cmat:C (72), pow (~40), + per element the centroid mapping (~600).
    "hex8_fbar": dict(config="2 mesh, TECH fbar (SURVEY 8f.1)", cell="HEX8", mat="MAT_COUP_NEOHOOKE", par=[200.0, 0.25],
                      matname="CoupNeoHooke (E=200, nu=0.25), F-bar", tech="TECH_FBAR",
                      flops=63.8e3 + 8 * (1152 + 576 + 288 + 72 + 40) + 600, bytes=2304.0,
                      kernel="k_hex8_eval_assemble<COUP_NEOHOOKE,STIFF,FBAR>", default_n=100,
                      metric="hex8 F-bar Neo-Hooke element evaluate+assemble (force + tangent stiffness into CSR), FP64"),
    # Section 8(f) row 1: EAS hex8 on the config-2 mesh.  Model = hex8 model + per Gauss point C*Mtilde (72 n),
    # Kaa (12 n^2), Kda (288 n), s (12 n) and per element the inverse (2 n^3), Kda*Kaa^-1 (48 n^2), the
    # condensed stiffness (1152 n) and force (48 n), n = 9 / 21 enhanced strain parameters.
    "hex8_eas_mild": dict(config="2 mesh, TECH eas_mild (SURVEY 8f.1)", cell="HEX8", mat="MAT_COUP_NEOHOOKE",
                          par=[200.0, 0.25], matname="CoupNeoHooke (E=200, nu=0.25), EAS mild", tech="TECH_EAS_MILD",
                          flops=63.8e3 + 8 * 4320 + 16146, bytes=2304.0 + 2 * 8 * 324,
                          kernel="k_hex8_eas_mma_eval_assemble<COUP_NEOHOOKE,STIFF,9> (FP64 tensor cores; option eas_mma=0: k_hex8_eas_eval_assemble)", default_n=100,
                          metric="hex8 EAS (mild) Neo-Hooke element evaluate+assemble (force + tangent stiffness into CSR), FP64"),
    "hex8_eas_full": dict(config="2 mesh, TECH eas_full (SURVEY 8f.1)", cell="HEX8", mat="MAT_COUP_NEOHOOKE",
                          par=[200.0, 0.25], matname="CoupNeoHooke (E=200, nu=0.25), EAS full", tech="TECH_EAS_FULL",
                          flops=63.8e3 + 8 * 13104 + 64890, bytes=2304.0 + 2 * 8 * 1008,
                          kernel="k_hex8_eas_mma_eval_assemble<COUP_NEOHOOKE,STIFF,21> (FP64 tensor cores; option eas_mma=0: k_hex8_eas_eval_assemble)", default_n=100,
                          metric="hex8 EAS (full) Neo-Hooke element evaluate+assemble (force + tangent stiffness into CSR), FP64"),
    # Section 8(f) row 4: history material on the device Gauss-point state store, KINEM linear hex8 on the config-2
    # mesh.  Model = hex8 model with M = ~150 flops (elastic predictor + return mapping + tangent); bytes add the
    # state store: 7 doubles read (last converged) + 7 written (current iterate) per Gauss point.
    "hex8_dp": dict(config="2 mesh, KINEM linear, MAT_Struct_DruckerPrager (SURVEY 8f.4)", cell="HEX8",
                    mat="MAT_DRUCKER_PRAGER", par=[10.0, 0.3, 0.5, 1e-12, 2e-3, 0.2, 0.25, 0.2, 50, 1.0],
                    matname="Drucker-Prager (E=10, nu=0.3, H=0.5, c=2e-3, eta=etabar=0.2, xi=0.25), KINEM linear",
                    tech="TECH_LINEAR", kinem="KINEM_LINEAR",
                    flops=8 * (62 * 64 + 228 * 8 + 1183 + 150), bytes=2304.0 + 8 * 14 * 8,
                    kernel="k_hex8_eval_assemble<DP,STIFF,LINEAR>", default_n=100,
                    metric="hex8 Drucker-Prager (history material) element evaluate+assemble (force + tangent stiffness into CSR), FP64"),
    "tet10": dict(config="4", cell="TET10", mat="MAT_ISO_VOL", par=[1.0, 1000.0],
                  matname="IsoNeoHooke mu=1 + VolSussmanBathe kappa=1000", flops=43.3e3, bytes=3300.0,
                  kernel="k_tet10_eval_assemble<ISO_VOL,STIFF>", default_n=87,
                  metric="tet10 iso/vol Neo-Hooke element evaluate+assemble (force + tangent stiffness into CSR), FP64"),
    # hex20 (20-node serendipity hexahedron) through the hex27 kernel with padded shape tables; model per Gauss
    # point 62 n^2 + 228 n + 1183 + M with n = 20 nodes, M = 1000, 27 points; bytes: 60 doubles of X, u,
    # 3600 + 60 K_e / f_e entries, 400 map entries
    "hex20": dict(config="3 mesh family, 20-node serendipity cell (not a BASELINE config)", cell="HEX20", mat="MAT_COUP_NEOHOOKE",
                  par=[200.0, 0.25], matname="CoupNeoHooke (E=200, nu=0.25)", flops=27 * (62 * 400 + 228 * 20 + 1183 + 1000),
                  bytes=8.0 * (120 + 3660) + 4 * 400, kernel="k_hex27_eval_assemble<COUP_NEOHOOKE,STIFF,0,PAD> (hex20 tables padded to 27 nodes)",
                  default_n=64, metric="hex20 Neo-Hooke element evaluate+assemble (force + tangent stiffness into CSR), FP64"),
    "hex27": dict(config="3", cell="HEX27", mat="MAT_SVK", par=[1000.0, 0.3], matname="St.Venant-Kirchhoff (E=1000, nu=0.3)",
                  flops=1.419e6, bytes=40e3, kernel="k_hex27_eval_assemble<SVK,STIFF>", default_n=80,
                  metric="hex27 St.Venant-Kirchhoff element evaluate+assemble (force + tangent stiffness into CSR), FP64"),
}


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the element kernel, PARSED from the newest
    committed ncu metrics CSV of this workload (profiles/r<NN>*_ncu_metrics_<workload>.csv, written by
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... --csv` around this very bench command; median
    over the captured launches).  Returns (bytes or None, file name or None)."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_ncu_metrics_{workload}.csv")))
    for path in reversed(files):
        per_launch = {}
        try:
            with open(path, newline="") as fh:
                rows = [r for r in csv.reader(l for l in fh if l.startswith('"'))]
        except OSError:
            continue
        if not rows:
            continue
        hdr = rows[0]
        try:
            ik, im, iv, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
        except ValueError:
            continue
        for r in rows[1:]:
            if "eval_assemble" not in r[ik] or not r[im].startswith("dram__bytes_"):
                continue
            if r[im] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                per_launch[r[ii]] = per_launch.get(r[ii], 0.0) + float(r[iv].replace(",", ""))
        if per_launch:
            return float(np.median(list(per_launch.values()))), os.path.basename(path)
    return None, None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.perf_counter()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.12] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def global_box(per_gpu_n, nranks, strong=False):
    """Weak scaling: every rank holds a per_gpu_n^3 element box (grid generator partition); strong scaling:
    the per_gpu_n^3 box is the whole mesh and is cut into nranks boxes."""
    b200 = importlib.import_module("4c_b200")
    mult = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(nranks)
    if mult is None:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    if strong:
        n = (per_gpu_n,) * 3
        return n, tuple(per_gpu_n / 100.0 for _ in range(3))
    n = tuple(per_gpu_n * m for m in mult)
    sub, _ = b200.partition.box_subdivisions(n, nranks)
    assert tuple(sub) == mult, (sub, mult)
    return n, tuple(m * per_gpu_n / 100.0 for m in mult)  # h = 0.01 as in config 2


def build_rank(wl, per_gpu_n, rank, nranks, strong=False):
    """Rank-local mesh + displacement state of the workload; returns (rank mesh, u on column nodes, label)."""
    b200 = importlib.import_module("4c_b200")
    cell = getattr(b200.mesh, wl["cell"])
    if wl["cell"] == "TET10":
        if nranks != 1:
            raise SystemExit("the tet10 workload (config 4) is a single-GPU configuration")
        X, conn = b200.mesh.unstructured_tet10(per_gpu_n)
        rm = b200.partition.serial_mesh(cell, X, conn)
        u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / per_gpu_n)
        return rm, u, f"unstructured tet10 ({per_gpu_n}^3 hex cells x 6, jittered, randomly renumbered)"
    if wl["cell"] == "HEX20":
        if nranks != 1:
            raise SystemExit("the hex20 workload is a single-GPU configuration")
        X, conn = b200.mesh.structured_hex20((per_gpu_n,) * 3, box=(per_gpu_n / 100.0,) * 3)
        rm = b200.partition.serial_mesh(cell, X, conn)
        u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.01)
        return rm, u, f"structured hex20 cube {per_gpu_n}^3"
    n, box = global_box(per_gpu_n, nranks, strong)
    rm = b200.partition.box_partition(cell, n, rank, nranks, box=box)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.01)
    return rm, u, f"structured {wl['cell'].lower()} cube {n[0]}x{n[1]}x{n[2]}"


# ------------------------------------------------------------------------------------------------
# CPU arms (the only place outside tests/ and smoke() that executes oracle/)
# ------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, wl, n_sample):
        b200 = importlib.import_module("4c_b200")
        from oracle import oracle as orc
        self.orc, self.wl = orc, wl
        self.kind = "reference" if orc.ref_lib() is not None else "port"
        self.impl = "ref" if self.kind == "reference" else "oracle"
        self.ct, self.mk = getattr(orc, wl["cell"]), getattr(orc, wl["mat"])
        self.tech = getattr(orc, wl.get("tech", "TECH_NONE"))
        if self.tech != orc.TECH_NONE or wl["cell"] == "HEX20":
            # oracle/_ref's assembled path covers the displacement-based formulation of hex8 / hex27 / tet10 only
            self.kind, self.impl = "port", "oracle"
        self.hist = None
        if wl["cell"] == "TET10":
            X, conn = b200.mesh.unstructured_tet10(n_sample)
            self.rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
            self.u = b200.mesh.displacement_field(self.rm.X, self.rm.node_gid, 1.0, 1.0 / n_sample)
        elif wl["cell"] == "HEX20":
            X, conn = b200.mesh.structured_hex20((n_sample,) * 3, box=(n_sample / 100.0,) * 3)
            self.rm = b200.partition.serial_mesh(b200.mesh.HEX20, X, conn)
            self.u = b200.mesh.displacement_field(self.rm.X, self.rm.node_gid, 1.0, 0.01)
        else:
            n = (n_sample,) * 3
            self.rm = b200.partition.box_partition(getattr(b200.mesh, wl["cell"]), n, 0, 1,
                                                   box=tuple(x / 100.0 for x in n))
            self.u = b200.mesh.displacement_field(self.rm.X, self.rm.node_gid, 1.0, 0.01)
        self.graph = orc.build_graph(self.ct, self.rm.conn, self.rm.node_row_lid, self.rm.node_col_lid,
                                     self.rm.n_row_nodes)
        self.n_sample = n_sample
        if self.tech in (orc.TECH_EAS_MILD, orc.TECH_EAS_FULL):
            self.hist = orc.eas_new_history(self.tech, self.rm.n_col_eles)
        if wl["mat"] == "MAT_DRUCKER_PRAGER":
            self.hist = np.zeros((self.rm.n_col_eles, 8, orc.DP_HIST))

    def step(self, threads):
        if self.hist is not None:
            threads = 1  # the oracle's history-carrying formulations are evaluated by one thread
        rm, orc = self.rm, self.orc
        t = time.perf_counter()
        _, _, nbad = orc.evaluate_assemble(self.ct, self.mk, self.wl["par"], rm.conn, rm.X.ravel(), self.u.ravel(),
                                           rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, self.graph[0], self.graph[1],
                                           nthreads=threads, impl=self.impl, tech=self.tech, hist=self.hist)
        dt = time.perf_counter() - t
        if nbad:  # a CPU arm that bailed out of its elements is not a baseline
            raise SystemExit(f"CPU arm ({self.impl}): {nbad} elements with a non-positive Jacobian determinant")
        return dt

    def describe(self):
        return (f"{self.rm.n_col_eles} {self.wl['cell'].lower()} {self.wl['matname']} elements of the same mesh "
                f"family (edge parameter {self.n_sample}, same displacement recipe), zero + evaluate + "
                f"owner-computes assembly")


def cpu_threads(wl, threads):
    """History-carrying formulations (EAS) run on one thread in the oracle."""
    return 1 if ("EAS" in wl.get("tech", "") or wl["mat"] == "MAT_DRUCKER_PRAGER") else threads


def cpu_sample_size(wl, threads, target_s):
    """Mesh edge parameter so that one CPU step takes about target_s (hex8: 33k elements/s/thread with
    oracle/_ref measured; scaled by the flop model for the other cells)."""
    threads = cpu_threads(wl, threads)
    rate = {"HEX8": 30e3, "TET10": 7e3, "HEX27": 1.3e3, "HEX20": 2.2e3}[wl["cell"]]  # elements/s/thread, measured
    rate *= 63.8e3 / wl["flops"] if wl["cell"] == "HEX8" else 1.0
    ne = rate * max(threads, 1) * 0.7 * target_s
    if wl["cell"] == "TET10":
        ne /= 6.0
    return int(max(4, min(wl["default_n"], round(ne ** (1.0 / 3.0)))))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    wl = WORKLOADS[args.workload]
    METRIC = wl["metric"]
    arm = CpuArm(wl, cpu_sample_size(wl, threads, 1.0))
    for _ in range(args.warmup):
        arm.step(threads)
    ts = [arm.step(threads) for _ in range(args.steps)]
    tot = sum(ts)
    val = arm.rm.n_col_eles * args.steps / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BASELINE config {wl['config']} ({wl['cell'].lower()}, {wl['matname']}); "
                               "each step is a bounded sample: " + arm.describe(),
                   "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cpu_threads(wl, threads), "kind": arm.kind,
                         "sample": arm.describe()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
PARITY_WORKLOADS = ("hex8", "hex8_fbar", "tet10", "hex27", "hex20")


def parity_record(args, wl, ctx, rm, u, world, dist, torch, n_random=150, n_interface=50):
    """Driver-visible correctness at every N: AFTER the timed regions every rank compares the CSR rows of ~200
    owned nodes (random ones plus nodes of elements that hold ghost nodes, i.e. rows fed by halo-imported
    displacements) of the K resident from its last evaluate with the CPU oracle evaluated on the elements
    touching those nodes (oracle/sampled_rows.py; the oracle is the checker, never timed)."""
    if args.no_parity:
        return {"ok": None, "skipped": "--no-parity"}
    if args.workload not in PARITY_WORKLOADS:
        return {"ok": None, "skipped": "history-carrying formulation: covered by tests/test_gpu_post.py"}
    from oracle import oracle as orc
    from oracle import sampled_rows
    b200 = importlib.import_module("4c_b200")
    capi = b200.capi
    h_u = np.ascontiguousarray(u[:rm.n_row_nodes].ravel())
    ctx.evaluate(capi.ACTION_NLNSTIFF, h_u, None, None)   # K resident for exactly this state
    nodes = sampled_rows.pick_nodes(rm, n_random, n_interface if world > 1 else 0, seed=17 + rm.rank)
    tech = getattr(orc, wl.get("tech", "TECH_NONE"))
    r = sampled_rows.check(ctx, rm, getattr(orc, wl["cell"]), getattr(orc, wl["mat"]), wl["par"], u, nodes,
                           rtol=1e-12, tech=tech, nthreads=min(8, host_threads()))
    v = torch.tensor([0.0 if r["ok"] else 1.0, r["worst_blockrow_ratio"], r["worst_entry_rel"], r["worst_f_ratio"],
                      float(r["nodes"]), float(r["entries"])], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = v.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = v.clone()
        dist.all_reduce(sm)
        v = torch.cat([mx[:4], sm[4:]])
    v = v.tolist()
    return {"ok": v[0] == 0.0, "worst": v[1], "worst_entry_rel": v[2], "worst_f": v[3], "rtol": 1e-12,
            "norm": "worst = max |a-b| / (rtol * max-abs of the node's 3 CSR rows) (SURVEY 8d, <= 1 passes); "
                    "worst_entry_rel = max |a-b| / max(|a|,|b|) over entries above 1e-6 of that scale",
            "rows_checked": int(3 * v[4]), "entries_checked": int(v[5]), "ranks_checked": world,
            "pattern": "bit-exact colind/rowptr of the sampled rows vs the oracle's FillComplete restatement",
            "ghost_fed_rows": world > 1}


def run_config5(args, wl, rank, world, local, barrier, dist, torch):
    """BASELINE config 5: 400^3 = 64 M hex8 Neo-Hooke elements on 8 B200 (200^3 row elements per rank, nnz/rank
    1.97e9 > 2^31), device-resident timing + sampled-row parity.  No e2e leg: the host staging of 8 x 15.8 GB of
    CSR values is not a sensible benchmark of this box's host."""
    b200 = importlib.import_module("4c_b200")
    capi = b200.capi
    rm, u, label = build_rank(wl, 200, rank, world, False)
    asm = b200.DeviceAssembler(rm, getattr(capi, wl["mat"]), wl["par"], device=local)
    ctx = asm.ctx
    uid = [ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init_nccl(uid[0])
    h_u = np.ascontiguousarray(u[:rm.n_row_nodes].ravel())
    ctx.evaluate(capi.ACTION_NLNSTIFF, h_u, None, None)
    stream = torch.cuda.ExternalStream(ctx.device_ptrs()["stream"], device=torch.device("cuda", local))
    for _ in range(args.warmup):
        ctx.evaluate_async(capi.ACTION_NLNSTIFF)
    ctx.synchronize()
    ctx.kernel_time(reset=True)
    steps = max(3, min(args.steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        ctx.evaluate_async(capi.ACTION_NLNSTIFF)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ctx.synchronize()
    ksum, kcnt = ctx.kernel_time(reset=True)
    info = ctx.last_evaluate_info()
    parity = parity_record(args, wl, ctx, rm, u, world, dist, torch)
    tot = torch.tensor([rm.n_row_eles, rm.n_col_eles], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    out = {"workload": f"BASELINE config 5: {label}, {wl['matname']}, {int(tot[0])} elements, nnz/rank={asm.nnz}",
           "value": float(tot[0]) * steps / (float(ms) * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
           "ms_per_step": float(ms) / steps, "kernel_ms_rank0": ksum / max(kcnt, 1),
           "halo_overlapped": bool(info["halo_overlapped"]), "ghost_elements_rank0": info["ghost_elements"],
           "parity": parity}
    asm.close()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    b200 = importlib.import_module("4c_b200")
    capi = b200.capi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                             "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl = WORKLOADS[args.workload]
    METRIC = wl["metric"]
    FLOPS_PER_ELE, BYTES_PER_ELE = wl["flops"], wl["bytes"]
    per_gpu_n = args.per_gpu_n or wl["default_n"]
    rm, u, mesh_label = build_rank(wl, per_gpu_n, rank, world, args.strong)
    asm = b200.DeviceAssembler(rm, getattr(capi, wl["mat"]), wl["par"], device=local,
                               tech=getattr(capi, wl.get("tech", "TECH_NONE")) if "kinem" not in wl else capi.TECH_NONE,
                               kinem=getattr(capi, wl.get("kinem", "KINEM_NONLINEAR_TOTLAG")),
                               options={kv.split("=")[0]: float(kv.split("=")[1]) for kv in args.option or []})
    ctx = asm.ctx
    if world > 1:
        uid = [ctx.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init_nccl(uid[0])
    n_rows, nnz = asm.n_rows, asm.nnz
    ptr = ctx.device_ptrs()
    stream = torch.cuda.ExternalStream(ptr["stream"], device=torch.device("cuda", local))
    u_row = np.ascontiguousarray(u[:rm.n_row_nodes].ravel())

    # pinned host staging of the reference-facing call (u in; K values, f out)
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except (OSError, IndexError, ValueError):
        avail = 1 << 62
    if 8 * (nnz + 2 * n_rows) * world > 0.5 * avail:
        raise SystemExit(f"host staging of {8 * nnz * world / 1e9:.1f} GB of CSR values does not fit the host "
                         f"({avail / 1e9:.1f} GB available): choose a smaller --per-gpu-n")
    # pinned staging on the NUMA node of this rank's GPU (b200s_host_alloc_near; --host-alloc plain = cudaMallocHost)
    hb, numa_node = {}, C.c_int(-1)
    for k, sz in (("u", n_rows), ("K", nnz), ("f", n_rows)):
        hb[k] = (capi.load().b200s_host_alloc_near(local, 8 * sz, C.byref(numa_node)) if args.host_alloc == "near"
                 else capi.load().b200s_host_alloc(8 * sz))
    if not all(hb.values()):
        raise SystemExit("pinned host allocation failed")
    h_u = np.ctypeslib.as_array((capi.C.c_double * n_rows).from_address(hb["u"]))
    h_u[:] = u_row
    # resident input for the device-timed arm
    ctx.evaluate(capi.ACTION_NLNSTIFF, h_u, None, None)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        own = e0.elapsed_time(e1)
        ms = torch.tensor([own], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        timed.own_ms = own
        return float(ms.item())

    def step_dev():
        ctx.evaluate_async(capi.ACTION_NLNSTIFF)

    def step_e2e():
        ctx.evaluate(capi.ACTION_NLNSTIFF, hb["u"], hb["K"], hb["f"])

    sampler = ClockSampler(local) if rank == 0 else None  # covers the device arm AND the e2e arm
    # ---- device-resident arm ----------------------------------------------------------------
    for _ in range(args.warmup):
        step_dev()
    ctx.synchronize()
    ctx.kernel_time(reset=True)
    l0 = ctx.launch_count()
    t0 = sampler.mark() if sampler else 0
    ms_dev = timed(step_dev, args.steps)
    _, err = ctx.synchronize()
    launches = ctx.launch_count() - l0
    ksum, kcnt = ctx.kernel_time(reset=True)
    tim = ctx.timers()
    # ---- end-to-end arm (host buffers through the blocking C-ABI call) -------------------
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, e2e_steps)
    e2e_info = ctx.last_evaluate_info()
    # per-rank device->host rate of the K values during the e2e steps (every rank's own clock)
    own_gbs = torch.tensor([8.0 * (nnz + n_rows) * e2e_steps / (timed.own_ms * 1e-3) / 1e9, float(numa_node.value)],
                           dtype=torch.float64, device="cuda")
    all_gbs = [torch.zeros_like(own_gbs) for _ in range(world)]
    if world > 1:
        dist.all_gather(all_gbs, own_gbs)
    else:
        all_gbs = [own_gbs]
    # ---- the step after the path with K resident on the device -----------------
    # e2e_resident: same blocking call, host u in / host f out, K stays in the context's CSR for
    # b200s_apply_dirichlet / b200s_multiply / a device solver; multiply: y = K x on that CSR.
    def step_res():
        ctx.evaluate(capi.ACTION_NLNSTIFF, hb["u"], None, hb["f"])

    for _ in range(2):
        step_res()
    ms_res = timed(step_res, e2e_steps)
    h_f = np.ctypeslib.as_array((capi.C.c_double * n_rows).from_address(hb["f"]))
    mult_ms = []
    for _ in range(2 + 5):
        ctx.multiply(h_f)
        mult_ms.append(ctx.last_multiply_ms())
    mult_ms = float(np.median(mult_ms[2:]))
    t1 = sampler.mark() if sampler else 0
    clocks = sampler.stop(t0, t1) if sampler else None

    parity = parity_record(args, wl, ctx, rm, u, world, dist if world > 1 else None, torch)
    halo_info = None
    if world > 1:
        ctx.evaluate_async(capi.ACTION_NLNSTIFF)
        ctx.synchronize()
        halo_info = ctx.last_evaluate_info()
    tot_eles = torch.tensor([rm.n_row_eles, rm.n_col_eles], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot_eles)
    row_eles, col_eles = float(tot_eles[0]), float(tot_eles[1])
    value = row_eles * args.steps / (ms_dev * 1e-3)
    e2e = row_eles * e2e_steps / (ms_e2e * 1e-3)
    e2e_res = row_eles * e2e_steps / (ms_res * 1e-3)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        dfma_tf, fp64_mhz = capi.measure_fp64_peak(local)
        dmma_tf = capi.microbench(1, 0, local)
        fp64_tf = max(dfma_tf, dmma_tf)   # the FP64 pipe's peak is whichever instruction stream reaches more
        k_ms = ksum / max(kcnt, 1)
        traffic, tfile = (args.traffic_bytes, "--traffic-bytes") if args.traffic_bytes is not None else (
            ncu_traffic(args.workload) if per_gpu_n == wl["default_n"] and world == 1 else (None, None))
        traffic_src = (f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch (median), parsed from profiles/{tfile}"
                       if traffic is not None else "no committed ncu capture for this workload/size")
        eps_kernel = rm.n_col_eles / (k_ms * 1e-3)      # elements this rank's kernel evaluates per second
        ach_tf = FLOPS_PER_ELE * eps_kernel / 1e12
        ach_gbs = BYTES_PER_ELE * eps_kernel / 1e9
        roof = {
            "bound": "fp64", "kernel": wl["kernel"],
            "achieved": ach_tf, "peak": fp64_tf, "unit": "TFLOP/s", "frac": ach_tf / fp64_tf,
            "peak_source": f"max of two microbenchmarks run in this process: DFMA loop {dfma_tf:.1f} TFLOP/s, DMMA "
                           f"m8n8k4 loop {dmma_tf:.1f} TFLOP/s (148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2); "
                           "MEASURED_PEAKS.json holds no FP64 figure",
            "model": f"SURVEY 8(d): {FLOPS_PER_ELE / 1e3:.1f} kflop and {BYTES_PER_ELE:.0f} B per {wl['cell'].lower()} "
                     "element (algorithmic, full non-symmetric K_e)",
            "kernel_ms": k_ms, "kernel_launches_timed": kcnt,
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": BYTES_PER_ELE * rm.n_col_eles,
            "hbm": {"achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach_gbs / peaks["hbm_gbs"], "peak_source": peak_src},
        }
        cpu = None
        if world == 1 and not args.no_cpu:
            threads = host_threads()
            arm = CpuArm(wl, cpu_sample_size(wl, threads, 3.0))
            arm.step(threads)
            ts = [arm.step(threads) for _ in range(3)]
            cpu = {"value": arm.rm.n_col_eles / float(np.median(ts)), "unit": UNIT, "cores": cpu_threads(wl, threads),
                   "kind": arm.kind, "sample": arm.describe() + ", median of 3 after 1 warm-up"}
            arm1 = CpuArm(wl, cpu_sample_size(wl, 1, 2.0))      # also P = 1
            arm1.step(1)
            cpu["value_1_core"] = arm1.rm.n_col_eles / min(arm1.step(1) for _ in range(2))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"BASELINE config {wl['config'] if world == 1 else '5-style weak'}: {mesh_label}, "
                            f"{wl['matname']}, {int(row_eles)} elements ({rm.n_row_eles} on rank 0), "
                            f"struct_calc_nlnstiff, nnz/rank={nnz}",
                "partition": f"{world} box(es), owner-computes rows, {int(col_eles - row_eles)} ghost elements evaluated redundantly",
                "l2": f"working set per step {8 * nnz / 1e6:.0f} MB of CSR values + {rm.n_col_eles * rm.conn.shape[1] ** 2 * 4 / 1e6:.0f} MB "
                      f"scatter map > L2 ({L2_BYTES / 1e6:.0f} MB): inputs larger than L2, no explicit flush",
                "timers_last_step_ms": {k: round(float(v), 4) for k, v in tim.items()},
            },
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * n_rows * world,
                    "d2h_bytes_per_step": 8 * (nnz + n_rows) * world, "ms_per_step": ms_e2e / e2e_steps,
                    "steps": e2e_steps, "api": "b200s_evaluate(ctx, NLNSTIFF, u_host, K_values_host, f_host) with pinned host buffers",
                    "pipeline": {"row_phases": e2e_info["row_phases"], "chunks_overlapped": e2e_info["pipeline_chunks"],
                                 "kernel_launches_per_step": e2e_info["kernel_launches"]},
                    "host_alloc": args.host_alloc,
                    "per_rank": [{"rank": r, "bytes_per_s_gb": round(float(g[0]), 2), "numa_node": int(g[1])}
                                 for r, g in enumerate(all_gbs)]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "error_info": {"n_neg_jacobian": err.n_neg_jacobian, "min_detJ": err.min_detJ, "min_detF": err.min_detF},
        }
        # node-blocked CSR product: 8 B per value + one 4 B column id per 3x3 block + x and y once
        mult_bytes = nnz * (8.0 + 4.0 / 9.0) + 8.0 * (3 * rm.n_col_nodes + n_rows)
        line["after_path"] = {
            "e2e_resident": {"value": e2e_res, "unit": UNIT, "ms_per_step": ms_res / e2e_steps, "steps": e2e_steps,
                             "h2d_bytes_per_step": 8 * n_rows * world, "d2h_bytes_per_step": 8 * n_rows * world,
                             "api": "b200s_evaluate(ctx, NLNSTIFF, u_host, NULL, f_host): K stays in the device CSR "
                                    "for b200s_apply_dirichlet / b200s_multiply / b200s_solve_cg"},
            "multiply": {"kernel": "k_spmv_node3 (+ column-map import of x)", "ms": mult_ms, "rank": 0,
                         "roofline": {"bound": "hbm", "achieved": mult_bytes / (mult_ms * 1e-3) / 1e9,
                                      "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                      "frac": mult_bytes / (mult_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                      "bytes_model": "nnz*(8 + 4/9) + 8*(n_cols + n_rows)", "peak_source": peak_src}},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        line["parity"] = parity
        if halo_info:
            line["config"]["halo"] = {"overlapped_with_interior_elements": bool(halo_info["halo_overlapped"]),
                                      "ghost_elements_rank0": halo_info["ghost_elements"],
                                      "kernel_launches_per_step": halo_info["kernel_launches"]}
    for p in hb.values():
        capi.load().b200s_host_free(p)
    asm.close()
    # BASELINE config 5 proper (64 M hex8 elements on 8 GPUs = 200^3 per rank): an extra device-resident run
    if world == 8 and args.workload == "hex8" and not args.strong and not args.per_gpu_n and not args.no_config5:
        c5 = run_config5(args, wl, rank, world, local, barrier, dist, torch)
        if rank == 0:
            line["config5"] = c5
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--option", action="append", help="library option key=value (b200s_set_option), repeatable")
    ap.add_argument("--workload", default="hex8", choices=sorted(WORKLOADS),
                    help="hex8 = BASELINE config 2 (default, the headline), tet10 = config 4, hex27 = config 3")
    ap.add_argument("--per-gpu-n", type=int, default=0,
                    help="edge parameter of each rank's mesh (default: the config's size; hex8 100 -> 1M elements)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: --per-gpu-n is the edge of the WHOLE mesh, cut into N boxes")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled-row parity record")
    ap.add_argument("--no-config5", action="store_true", help="at --gpus 8: skip the extra 200^3-per-rank run")
    ap.add_argument("--host-alloc", default="near", choices=["near", "plain"],
                    help="pinned staging of the e2e leg: on the GPU's NUMA node (default) or plain cudaMallocHost")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch of the element kernel from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
