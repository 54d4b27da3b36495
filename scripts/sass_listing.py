"""Full SASS of the hot kernels from the built object, with a mnemonic census on top:
    python scripts/sass_listing.py [round tag, default r02]
-> profiles/<tag>_sass_<kernel>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
obj = os.path.join(ROOT, "mimick_b200", "_build", "mmk_device.o")
kernels = {
    "k_emit_w_2bit": "_ZN6mmkdev8k_emit_wILi2ELb0EEEv8EmitArgs",
    "k_emit_w_2bit_errmodel": "_ZN6mmkdev8k_emit_wILi2ELb1EEEv8EmitArgs",
    "k_place_block": "_ZN6mmkdev13k_place_blockENS_9PlaceArgsE",
    "k_bgzf": None,
    "k_insert_draws": None,
    "k_place_chain": None,
    "k_plan_molecules": None,
}
syms = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
names = sorted(set(re.findall(r"\.text\.(_ZN6mmkdev\w+)", syms)))
for k, m in kernels.items():
    if m is None:
        cands = [n for n in names if re.search(r"\d+" + k + r"E", n)]
        m = cands[0] if cands else None
    if m is None or m not in names:
        print("not found:", k)
        continue
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", m, obj], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for line in sass.split("\n"):
        mm = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if mm:
            ops[mm.group(1)] += 1
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_{k}.txt")
    with open(path, "w") as fh:
        fh.write(f"# {m}\n# nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo; {sum(ops.values())} instructions\n# mnemonic census:\n")
        for op, c in ops.most_common():
            fh.write(f"#   {c:6d}  {op}\n")
        for line in sass.split("\n"):  # drop the instruction encodings: half the lines and most of the bytes
            line = re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line)
            if line.strip():
                fh.write(line.rstrip() + "\n")
    print(path, sum(ops.values()), "instructions")
