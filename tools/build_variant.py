"""Development tool: build an A/B variant of the library with extra preprocessor defines.
    python tools/build_variant.py <tag> [-DNAME=value ...]   ->  tools/_libs/<tag>.so
(time it next to the product build with tools/lib_probe.py; the product itself is built by __graft_entry__.build())."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

tag, defines = sys.argv[1], sys.argv[2:]
nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
obj_dir = os.path.join(ROOT, "build", f"obj_{tag}")
os.makedirs(obj_dir, exist_ok=True)
os.makedirs(os.path.join(ROOT, "tools", "_libs"), exist_ok=True)
src = os.path.join(ROOT, "4c_b200", "csrc", "b200_solid.cu")
flags = [f for f in g.NVCC_FLAGS if f != "-shared"] + ["-diag-suppress", "177"] + defines
procs, objs = [], []
for part in range(5):
    obj = os.path.join(obj_dir, f"part{part}.o")
    objs.append(obj)
    procs.append(subprocess.Popen([nvcc] + flags + [f"-DB200S_PART={part}", "-c", "-o", obj, src], cwd=ROOT))
assert all(p.wait() == 0 for p in procs)
out = os.path.join(ROOT, "tools", "_libs", f"{tag}.so")
subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs, cwd=ROOT)
print(out)
