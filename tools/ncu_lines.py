"""Map ncu SASS-level samples back to CUDA source lines without a GUI.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-mangled-substring> [lib.so]"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "liorf_b200", "lib", "libliorf_b200.so")
tmp = tempfile.mkdtemp()
if lib.endswith(".cubin"):
    cubin_path = lib
else:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin_path = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin_path], capture_output=True, text=True).stdout.splitlines()
# instruction index -> (file, line) for the kernel section
inside = False; cur = ("?", 0); lines = []
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
his = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Address"]
sec = int(os.environ.get("NCU_SECTION", "0"))            # report with several launches: pick the launch (0-based)
hi = his[sec]
rows = rows[:his[sec + 1] - 2] if sec + 1 < len(his) else rows
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0; toti = 0
for k, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr) or not r[ci["# Samples"]].isdigit():
        continue
    s = int(r[ci["# Samples"]]); ie = int(r[ci["Instructions Executed"]])
    key = lines[k] if k < len(lines) else ("?", -1)
    agg[key][0] += s; agg[key][1] += ie
    for c in stall_cols:
        v = int(r[ci[c]] or 0)
        if v:
            agg[key][2][c] += v
    tot += s; toti += ie
print(f"total samples {tot}, warp instructions {toti}, sass instrs {len(rows) - hi - 1}, disasm instrs {len(lines)}")
src_cache = {}
by_inst = os.environ.get("NCU_SORT", "samples") == "inst"            # NCU_SORT=inst: order by executed instructions instead of stall samples
for key, (s, ie, st) in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_inst else 0])[:int(os.environ.get("NCU_TOP", "40"))]:
    f, l = key
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "liorf_b200", "csrc", f)
    if f not in src_cache:
        try:
            src_cache[f] = open(path).read().splitlines()
        except OSError:
            src_cache[f] = []
    text = src_cache[f][l - 1].strip()[:90] if 0 < l <= len(src_cache[f]) else ""
    top = ",".join(f"{c[6:]}:{v}" for c, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * s / max(tot, 1):5.1f}% smp {100 * ie / max(toti, 1):5.1f}% inst  {f}:{l:<4} [{top}]  {text}")
