"""BASELINE.json configs[0] and configs[2]: duct 24x8x8 (1536 elements), chi = 1000 in the lowperm box; lx = 6 and 8,
GLL-grid and dealiased operator.  Device-resident step time vs the CPU oracle on the same mesh (all host threads).
usage: python tools/duct_bench.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import neko_top_b200  # noqa: E402,F401
from neko_top_b200 import operators as ops, sem, workloads  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402  (reported CPU baseline only)

orc.build()
for lx in (6, 8):
    brick = workloads.config_duct(lx)
    sp = sem.Space(lx)
    x, y, z = workloads.coords(brick, "cuda")
    keys = workloads.node_keys(brick, "cuda")
    G, _, B = sem.geometric_factors(x, y, z, sp)
    fl = workloads.make_fields(brick, x, y, z, keys)
    chi = workloads.brinkman_zone_chi(x, y, z)
    flat = lambda a: a.reshape(-1).contiguous()
    Gf, Bf, v, ub, chif = [flat(g) for g in G], flat(B), [flat(a) for a in fl.v], [flat(a) for a in fl.ub], flat(chi)
    n = brick.n
    op = ops.fused_adjoint_rhs_t(ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), brick.nelv, Gf, Bf))
    op.gs.init(flat(keys))
    f = [torch.empty(n, device="cuda", dtype=torch.float64) for _ in range(3)]
    sens = torch.empty(n, device="cuda", dtype=torch.float64)
    c = lambda t: t.cpu().numpy()
    cid, nc = orc.gs_classes(c(flat(keys)))
    for dealias in (False, True):
        op.set_dealias(dealias)
        for _ in range(5):
            op.step(v, ub, f, chi=chif, sens=sens)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200
        e0.record()
        for _ in range(reps):
            op.step(v, ub, f, chi=chif, sens=sens)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        args = ([c(a) for a in v], [c(a) for a in ub], lx, brick.nelv, sp.dx, sp.wx, [c(g) for g in Gf], c(Bf))
        kw = dict(chi=c(chif), lxd=(3 * lx // 2 if dealias else 0))
        orc.adjoint_rhs(*args, **kw)
        t0 = time.perf_counter()
        nrep = 5
        for _ in range(nrep):
            fo, so, _ = orc.adjoint_rhs(*args, **kw)
            fo = [orc.gs_add(a, cid, nc) for a in fo]
        cpu_ms = (time.perf_counter() - t0) / nrep * 1e3
        print(json.dumps({"config": "configs[0]" if lx == 6 else "configs[2]", "lx": lx, "dealias": dealias, "dof": n,
                          "gpu_ms_per_step": ms, "gpu_gdof_s": n / ms / 1e6, "cpu_oracle_ms": cpu_ms,
                          "cpu_threads": orc.num_threads(), "cpu_gdof_s": n / cpu_ms / 1e6}), flush=True)
    op.free()
