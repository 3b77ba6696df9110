"""BASELINE.json configs[4]: polynomial-order sweep lx = 5..10 at ~1e8 DOF on one GPU (device-resident step).
usage: python tools/lxsweep.py [target_dof]   -> one JSON line per lx"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import neko_top_b200  # noqa: E402,F401
from neko_top_b200 import operators as ops, sem, workloads  # noqa: E402

target = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0e8
lxs = [int(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else [5, 6, 7, 8, 9, 10]
for lx in lxs:
    ne = max(2, round((target / lx ** 3) ** (1.0 / 3.0)))
    brick = workloads.config_box(ne, lx)
    sp = sem.Space(lx)
    x, y, z = workloads.coords(brick, "cuda")
    keys = workloads.node_keys(brick, "cuda")
    G, jac, B = sem.geometric_factors(x, y, z, sp, chunk=8192)
    fl = workloads.make_fields(brick, x, y, z, keys)
    del x, y, z, jac
    flat = lambda a: a.reshape(-1).contiguous()
    G, B = [flat(g) for g in G], flat(B)
    v, ub, rho = [flat(a) for a in fl.v], [flat(a) for a in fl.ub], flat(fl.rho)
    del fl
    n = brick.n
    f = [torch.empty(n, device="cuda", dtype=torch.float64) for _ in range(3)]
    sens = torch.empty(n, device="cuda", dtype=torch.float64)
    op = ops.fused_adjoint_rhs_t(ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), brick.nelv, G, B))
    op.gs.init(flat(keys))
    del keys
    for _ in range(3):
        op.step(v, ub, f, rho=rho, sens=sens)
    torch.cuda.synchronize()
    op.enable_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        op.step(v, ub, f, rho=rho, sens=sens)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ek, gk, _ = op.get_timing()
    print(json.dumps({"cfg": os.environ.get("B200_ADJRHS_CFG", ""), "lx": lx, "ne": ne, "dof": n, "ms_per_step": ms, "gdof_s": n / ms / 1e6, "elem_ms": ek,
                      "gs_ms": gk, "elem_GBps": 168.0 * n / ek / 1e6,
                      "step_GBps_algorithmic": sem.algorithmic_bytes_per_dof(lx) * n / ms / 1e6}), flush=True)
    op.free()
    del G, B, v, ub, rho, f, sens, op
    torch.cuda.empty_cache()
