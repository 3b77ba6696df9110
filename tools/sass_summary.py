"""Per-kernel SASS opcode summary of libneko_top_b200.so (cuobjdump -sass): the evidence that the hot kernels are
hand-written sm_100a code -- DMMA (fp64 tensor-core mma.sync.m8n8k4.f64; tcgen05 has no fp64 type), UBLKCP / UBLKPF
(1-D TMA bulk copies / L2 prefetch), SYNCS (mbarrier), 128-bit LDG/STG/LDS.  Runs without a GPU.
usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "neko-top_b200", "libneko_top_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
kern, cur = collections.OrderedDict(), None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        kern[cur][m.group(1)] += 1
KEY = ["DMMA", "DFMA", "UBLKCP", "UBLKPF", "SYNCS", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64", "LDS.128",
       "LDS.64", "STS.128", "LDC", "BAR", "SHFL", "ATOMG", "RED"]
only = sys.argv[1:] or ["adjrhs_v3_kernel", "adjrhs_v2_kernel", "advop_mma_kernel", "advop_kernel", "gs_op_kernel", "gs_face_pass_kernel",
                        "helm_kernel", "deriv_kernel"]
print(f"# SASS opcode counts per kernel (static), {os.path.basename(so)}, sm_100a; cuobjdump -sass")
print("# " + "  ".join(KEY) + "  | total")
for name, cnt in kern.items():
    d = demangle(name)
    if not any(o in d for o in only):
        continue
    short = re.sub(r"\(.*", "", d.replace("b200::", "").replace("void ", ""))
    row = []
    for k in KEY:
        row.append(sum(v for op, v in cnt.items() if op == k or op.startswith(k + ".")))
    print(f"{short[:78]:78s} " + " ".join(f"{v:5d}" for v in row) + f" | {sum(cnt.values())}")
