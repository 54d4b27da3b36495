"""Attribute a kernel's executed instructions / stall samples to CUDA source lines.

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep 'k_emitILi2' [top]

ncu's CSV source page is SASS-only; nvdisasm -g gives the line of every SASS instruction of the
same cubin, so the two are zipped in instruction order.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, mangled = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
# NCU_OBJ: the object file of the build that was profiled (default: the current build)
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("NCU_OBJ", os.path.join(ROOT, "mimick_b200", "_build", "mmk_device.o"))], cwd=tmp,
               capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if re.match(r"\s*\.section\s+\.text\.\S*" + re.escape(mangled), l)][0]
lines, cur = [], None
for l in dis[start + 1:]:
    if re.match(r"\s*\.section", l):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+.*?;", l):
        lines.append(cur)
skip = os.environ.get("NCU_SKIP", "0")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if r and r[0].startswith("0x")][: len(lines)]
ie, ism = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg, smp = collections.Counter(), collections.Counter()
for cur, r in zip(lines, data):
    agg[cur] += int(r[ie])
    smp[cur] += int(r[ism])
tot, tsm = sum(agg.values()), max(sum(smp.values()), 1)
src = {}
print(f"instructions executed: {tot}  ({len(lines)} SASS instructions)")
for k, v in agg.most_common(top):
    f, ln = k if k else ("?", 0)
    if f not in src:
        p = os.path.join(ROOT, "mimick_b200", "csrc", f)
        if not os.path.exists(p):
            p = os.path.join(ROOT, "include", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
    print(f"{100 * v / tot:5.1f}% inst {100 * smp[k] / tsm:5.1f}% stall-samples  {f}:{ln}  {text}")
