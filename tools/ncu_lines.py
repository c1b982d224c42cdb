"""Development tool: join the SASS page of an ncu report (stall samples, executed instructions, shared-memory wavefronts
per instruction) with nvdisasm's line table of the same kernel and print the hottest source lines.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep build/obj/b200_solid_part2.o <mangled kernel substring> [n_elements] [top]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, sub = sys.argv[1:4]
nele = float(sys.argv[4]) if len(sys.argv) > 4 else 1e6
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
ins, cur, on = [], None, False
for ln in dis.splitlines():
    if ln.startswith("//-----"):
        if on and ins:
            break
        on = sub in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
assert len(data) == len(ins), (len(data), len(ins))
ia, isamp, iw = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot_st = collections.Counter()
for r, (li, txt) in zip(data, ins):
    a = agg[li]
    a[0] += int(r[ia] or 0)
    a[1] += int(r[isamp] or 0)
    a[2] += int(float(r[iw] or 0))
    for i in st:
        if r[i]:
            a[3][hdr[i][6:]] += int(r[i])
            tot_st[hdr[i][6:]] += int(r[i])
ti, ts, tw = (sum(a[k] for a in agg.values()) for k in range(3))
print(f"instructions / element {ti / nele:.0f}, shared wavefronts / element {tw / nele:.0f}, samples {ts}")
print("stalls %:", {k: round(100 * v / sum(tot_st.values()), 1) for k, v in tot_st.most_common(8)})
byfile = collections.Counter()
for (li, a) in agg.items():
    byfile[li[0] if li else None] += a[1]
print("by file %:", {k: round(100 * v / ts, 1) for k, v in byfile.most_common(8)})
for li, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{str(li):44s} inst/el {a[0] / nele:7.1f}  samples {100 * a[1] / ts:5.1f}%  shwf/el {a[2] / nele:6.1f}  {dict(a[3].most_common(2))}")
