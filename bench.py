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
                          kernel="k_hex8_eas_eval_assemble<COUP_NEOHOOKE,STIFF,9>", default_n=100,
                          metric="hex8 EAS (mild) Neo-Hooke element evaluate+assemble (force + tangent stiffness into CSR), FP64"),
    "hex8_eas_full": dict(config="2 mesh, TECH eas_full (SURVEY 8f.1)", cell="HEX8", mat="MAT_COUP_NEOHOOKE",
                          par=[200.0, 0.25], matname="CoupNeoHooke (E=200, nu=0.25), EAS full", tech="TECH_EAS_FULL",
                          flops=63.8e3 + 8 * 13104 + 64890, bytes=2304.0 + 2 * 8 * 1008,
                          kernel="k_hex8_eas_eval_assemble<COUP_NEOHOOKE,STIFF,21>", default_n=100,
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
    "hex27": dict(config="3", cell="HEX27", mat="MAT_SVK", par=[1000.0, 0.3], matname="St.Venant-Kirchhoff (E=1000, nu=0.3)",
                  flops=1.419e6, bytes=40e3, kernel="k_hex27_eval_assemble<SVK,STIFF>", default_n=80,
                  metric="hex27 St.Venant-Kirchhoff element evaluate+assemble (force + tangent stiffness into CSR), FP64"),
}


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the element kernel at the workload's default
# size, from the committed ncu captures profiles/r01e_ncu_metrics_<workload>.csv (same bench.py command)
NCU_TRAFFIC_BYTES = {"hex8": 2694710016 + 2248577536, "tet10": 20888573696 + 17315611904,
                     "hex27": 25082223616 + 24644916992, "hex8_eas_full": 3559638528 + 10256551424}
NCU_TRAFFIC_FILE = {"hex8": "r01h_ncu_metrics_hex8.csv", "hex8_eas_full": "r01h_ncu_metrics_hex8_eas_full.csv"}


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
        if self.tech != orc.TECH_NONE:  # oracle/_ref covers the displacement-based formulation only
            self.kind, self.impl = "port", "oracle"
        self.hist = None
        if wl["cell"] == "TET10":
            X, conn = b200.mesh.unstructured_tet10(n_sample)
            self.rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
            self.u = b200.mesh.displacement_field(self.rm.X, self.rm.node_gid, 1.0, 1.0 / n_sample)
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
        orc.evaluate_assemble(self.ct, self.mk, self.wl["par"], rm.conn, rm.X.ravel(), self.u.ravel(),
                              rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, self.graph[0], self.graph[1],
                              nthreads=threads, impl=self.impl, tech=self.tech, hist=self.hist)
        return time.perf_counter() - t

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
    rate = {"HEX8": 30e3, "TET10": 7e3, "HEX27": 1.3e3}[wl["cell"]]  # elements/s/thread, measured
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
                               kinem=getattr(capi, wl.get("kinem", "KINEM_NONLINEAR_TOTLAG")))
    ctx = asm.ctx
    for kv in args.option or []:          # library options, e.g. --option lazy_zero=0
        k, v = kv.split("=")
        ctx.set_option(k, float(v))
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
    hb = {k: capi.load().b200s_host_alloc(8 * sz) for k, sz in (("u", n_rows), ("K", nnz), ("f", n_rows))}
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
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
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

    tot_eles = torch.tensor([rm.n_row_eles, rm.n_col_eles], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot_eles)
    row_eles, col_eles = float(tot_eles[0]), float(tot_eles[1])
    value = row_eles * args.steps / (ms_dev * 1e-3)
    e2e = row_eles * e2e_steps / (ms_e2e * 1e-3)
    e2e_res = row_eles * e2e_steps / (ms_res * 1e-3)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        fp64_tf, fp64_mhz = capi.measure_fp64_peak(local)
        k_ms = ksum / max(kcnt, 1)
        eps_kernel = rm.n_col_eles / (k_ms * 1e-3)      # elements this rank's kernel evaluates per second
        ach_tf = FLOPS_PER_ELE * eps_kernel / 1e12
        ach_gbs = BYTES_PER_ELE * eps_kernel / 1e9
        roof = {
            "bound": "fp64", "kernel": wl["kernel"],
            "achieved": ach_tf, "peak": fp64_tf, "unit": "TFLOP/s", "frac": ach_tf / fp64_tf,
            "peak_source": "DFMA microbenchmark run in this process (b200s_measure_fp64_peak, implied "
                           f"{fp64_mhz:.0f} MHz); MEASURED_PEAKS.json holds no FP64 figure",
            "model": f"SURVEY 8(d): {FLOPS_PER_ELE / 1e3:.1f} kflop and {BYTES_PER_ELE:.0f} B per {wl['cell'].lower()} "
                     "element (algorithmic, full non-symmetric K_e)",
            "kernel_ms": k_ms, "kernel_launches_timed": kcnt,
            "traffic": args.traffic_bytes if args.traffic_bytes is not None else (
                NCU_TRAFFIC_BYTES.get(args.workload) if per_gpu_n == wl["default_n"] else None),
            "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/"
                              + NCU_TRAFFIC_FILE.get(args.workload, f"r01e_ncu_metrics_{args.workload}.csv"),
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
                    "steps": e2e_steps, "api": "b200s_evaluate(ctx, NLNSTIFF, u_host, K_values_host, f_host) with pinned host buffers"},
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
        print(json.dumps(line), flush=True)
    for p in hb.values():
        capi.load().b200s_host_free(p)
    asm.close()
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
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch of the element kernel from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
