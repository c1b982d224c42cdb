"""Development probe for the N>1 end-to-end collapse: concurrent device->host copies
from G GPUs of one box, one process per GPU, with the pinned staging allocated (a) by cudaMallocHost from
wherever the process happens to run and (b) on the GPU's NUMA node (b200s_host_alloc_near).  Also records the
box topology (nvidia-smi topo -m, NUMA nodes, affinity) and the host's own memory bandwidth with all threads
(what a host-side mirror of a symmetric K would be bounded by).  Prints JSON lines.

    python tools/d2h_probe.py [--gb 2] [--reps 5]
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(dev, ngroup, mode, nbytes, reps, bar, q, cpu_bind):
    import importlib
    capi = importlib.import_module("4c_b200").capi
    L = capi.load()
    rt = C.CDLL("libcudart.so.12")
    vp = C.c_void_p
    rt.cudaMalloc.argtypes = [C.POINTER(vp), C.c_size_t]
    rt.cudaMemcpyAsync.argtypes = [vp, vp, C.c_size_t, C.c_int, vp]
    rt.cudaEventCreate.argtypes = [C.POINTER(vp)]
    rt.cudaEventRecord.argtypes = [vp, vp]
    rt.cudaEventSynchronize.argtypes = [vp]
    rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), vp, vp]
    rt.cudaMemset.argtypes = [vp, C.c_int, C.c_size_t]
    node = C.c_int(-1)
    if cpu_bind:
        # run the process on the CPUs of the GPU's NUMA node as well
        try:
            n = int(open(f"/sys/bus/pci/devices/{pci_of(dev)}/numa_node").read())
            cpus = parse_cpulist(open(f"/sys/devices/system/node/node{n}/cpulist").read())
            os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
        except Exception:
            pass
    assert rt.cudaSetDevice(dev) == 0
    d = vp()
    assert rt.cudaMalloc(C.byref(d), nbytes) == 0
    rt.cudaMemset(d, 1, nbytes)
    h = L.b200s_host_alloc_near(dev, nbytes, C.byref(node)) if mode == "near" else L.b200s_host_alloc(nbytes)
    assert h
    e0, e1 = vp(), vp()
    rt.cudaEventCreate(C.byref(e0)), rt.cudaEventCreate(C.byref(e1))
    rt.cudaMemcpyAsync(h, d, nbytes, 2, None)
    rt.cudaDeviceSynchronize()
    best, tot = 0.0, 0.0
    for _ in range(reps):
        bar.wait()
        rt.cudaEventRecord(e0, None)
        rt.cudaMemcpyAsync(h, d, nbytes, 2, None)
        rt.cudaEventRecord(e1, None)
        rt.cudaEventSynchronize(e1)
        ms = C.c_float(0)
        rt.cudaEventElapsedTime(C.byref(ms), e0, e1)
        g = nbytes / (ms.value * 1e-3) / 1e9
        best, tot = max(best, g), tot + g
    # H2D as well (u goes that way; small, but shows the link's other direction)
    bar.wait()
    rt.cudaEventRecord(e0, None)
    rt.cudaMemcpyAsync(d, h, nbytes, 1, None)
    rt.cudaEventRecord(e1, None)
    rt.cudaEventSynchronize(e1)
    ms = C.c_float(0)
    rt.cudaEventElapsedTime(C.byref(ms), e0, e1)
    q.put({"dev": dev, "d2h_gbs_mean": round(tot / reps, 2), "d2h_gbs_best": round(best, 2),
           "h2d_gbs": round(nbytes / (ms.value * 1e-3) / 1e9, 2), "numa_node": node.value,
           "cpus": len(os.sched_getaffinity(0))})
    L.b200s_host_free(h)


def pci_of(dev):
    out = subprocess.run(["nvidia-smi", "-i", str(dev), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                         capture_output=True, text=True).stdout.strip().lower()
    return out[4:] if out.startswith("0000") and len(out) > 12 else out  # 00000000:17:00.0 -> 0000:17:00.0


def parse_cpulist(s):
    cpus = set()
    for part in s.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus |= set(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def host_bandwidth(nbytes=1 << 30):
    """memcpy bandwidth of the host with every available thread (numpy copies in a thread pool release the GIL)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    nt = len(os.sched_getaffinity(0))
    n = nbytes // 8 // nt
    src = [np.ones(n) for _ in range(nt)]
    dst = [np.empty(n) for _ in range(nt)]
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(lambda k: np.copyto(dst[k], src[k]), range(nt)))
        t = time.perf_counter()
        for _ in range(3):
            list(ex.map(lambda k: np.copyto(dst[k], src[k]), range(nt)))
        dt = (time.perf_counter() - t) / 3
    return {"threads": nt, "copy_gbs_read_plus_write": round(2 * 8 * n * nt / dt / 1e9, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=2.0)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    nbytes = int(args.gb * (1 << 30))
    ngpu = int(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU "))
    topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
    nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    info = {"gpus": ngpu, "numa_nodes": nodes, "affinity_cpus": len(os.sched_getaffinity(0)),
            "gpu_numa": {}, "node_cpus": {}}
    for d in range(ngpu):
        try:
            info["gpu_numa"][d] = int(open(f"/sys/bus/pci/devices/{pci_of(d)}/numa_node").read())
        except Exception as e:
            info["gpu_numa"][d] = str(e)[:60]
    for nd in nodes:
        try:
            info["node_cpus"][nd] = open(f"/sys/devices/system/node/{nd}/cpulist").read().strip()
        except Exception:
            pass
    print(json.dumps({"topology": info}), flush=True)
    print(topo, flush=True)
    print(json.dumps({"host_memcpy": host_bandwidth()}), flush=True)
    groups = [[0]]
    if ngpu >= 2:
        groups.append([0, 1])
    if ngpu >= 8:
        groups += [[0, 4], [0, 2, 4, 6], [0, 1, 2, 3], list(range(8))]
    elif ngpu >= 4:
        groups += [[0, 2], list(range(4))]
    ctx = mp.get_context("spawn")
    for grp in groups:
        for mode, bind in (("plain", False), ("near", False), ("near", True)):
            bar, q = ctx.Barrier(len(grp)), ctx.Queue()
            ps = [ctx.Process(target=worker, args=(d, len(grp), mode, nbytes, args.reps, bar, q, bind)) for d in grp]
            [p.start() for p in ps]
            res = sorted((q.get(timeout=300) for _ in ps), key=lambda r: r["dev"])
            [p.join() for p in ps]
            print(json.dumps({"group": grp, "host_alloc": mode, "cpu_bind": bind,
                              "sum_d2h_gbs": round(sum(r["d2h_gbs_mean"] for r in res), 1), "per_gpu": res}), flush=True)


if __name__ == "__main__":
    main()
