"""Static SASS opcode histogram of the hot kernels in libvgb200.so (what the staging path compiles to: STS vs
generic ST, UBLKCP = cp.async.bulk, PRMT = packed 16-bit index words, ATOMG = group tickets ...).
   python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vg_renderer_b200", "lib", "libvgb200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ["k_group_emitItLj1ELi1ELi1ELb0", "k_group_emitIjLj1ELi1ELi1ELb0", "k_group_emitItLj7ELi3ELi3ELb0", "k_group_countILj1ELi1ELi1", "k_flatten_slots",
        "k_fill_emitItLb0", "k_job_prep"]
arch = re.search(r"arch = (sm_\w+)", txt)
print(f"{os.path.basename(lib)}: {arch.group(1) if arch else '?'} (cuobjdump -sass)\n")
for fn in re.split(r"\n\s*Function : ", txt)[1:]:
    name = fn.split("\n", 1)[0].strip()
    if not any(w in name for w in want):
        continue
    ins = re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", fn, re.M)
    c = collections.Counter()
    for t in ins:
        t = t.strip().split()
        op = t[1] if t[0].startswith("@") else t[0]
        key = op if op.split(".")[0] in ("ST", "STS", "STG", "STL", "LDL", "LDS", "LDG", "LD", "ATOMG", "UBLKCP", "CCTL", "SHFL", "VOTE", "MATCH", "REDUX") else op.split(".")[0]
        c[key] += 1
    print(f"== {name}: {len(ins)} instructions")
    groups = {"shared stores (STS*)": sum(v for k, v in c.items() if k.startswith("STS")),
              "global stores (STG*)": sum(v for k, v in c.items() if k.startswith("STG")),
              "generic stores (ST.E*)": sum(v for k, v in c.items() if k.startswith("ST.")),
              "local memory (STL/LDL)": sum(v for k, v in c.items() if k.startswith("STL") or k.startswith("LDL")),
              "bulk async copies (UBLKCP)": sum(v for k, v in c.items() if k.startswith("UBLKCP")),
              "byte permutes (PRMT)": c.get("PRMT", 0), "shuffles (SHFL*)": sum(v for k, v in c.items() if k.startswith("SHFL")),
              "global atomics (ATOMG*)": sum(v for k, v in c.items() if k.startswith("ATOMG")),
              "L2 prefetches (CCTL*)": sum(v for k, v in c.items() if k.startswith("CCTL")),
              "tensor-core ops (UTC*MMA/HMMA)": sum(v for k, v in c.items() if "MMA" in k)}
    for k, v in groups.items():
        print(f"   {k:34s} {v}")
    print("   " + ", ".join(f"{k} {v}" for k, v in c.most_common(40)))
    print()
