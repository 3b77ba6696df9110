"""Device timing of the fine-grid advection operators (dealiased adjoint / linearised) on a box mesh.
usage: python tools/advop_bench.py [ne] [lx]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neko_top_b200  # noqa
from neko_top_b200 import operators as ops, sem, workloads

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 24
lx = int(sys.argv[2]) if len(sys.argv) > 2 else 8
brick = workloads.config_box(ne, lx)
sp = sem.Space(lx)
x, y, z = workloads.coords(brick, "cuda")
keys = workloads.node_keys(brick, "cuda")
G, _, B = sem.geometric_factors(x, y, z, sp, chunk=8192)
fl = workloads.make_fields(brick, x, y, z, keys)
flat = lambda a: a.reshape(-1).contiguous()
coef = ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), brick.nelv, [flat(g) for g in G], flat(B))
v, ub, rho = [flat(a) for a in fl.v], [flat(a) for a in fl.ub], flat(fl.rho)
n = brick.n
f = [torch.zeros(n, device="cuda", dtype=torch.float64) for _ in range(3)]
sens = torch.empty(n, device="cuda", dtype=torch.float64)
op = ops.fused_adjoint_rhs_t(coef)
adv = ops.adv_lin_dealias_b200_t(); adv.init(None, coef, op.handle)
lin = ops.adv_lin_b200_t(); lin.init(coef, op.handle)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

res = {"ne": ne, "lx": lx, "dof": n}
res["adjoint_dealias_ms"] = timeit(lambda: adv.compute_adjoint(*v, *ub, *f))
res["linear_dealias_ms"] = timeit(lambda: adv.compute_linear(*v, *ub, *f))
res["linear_gll_ms"] = timeit(lambda: lin.compute_linear(*v, *ub, *f))
res["fused_gll_ms"] = timeit(lambda: op.compute(v, ub, f, rho=rho, sens=sens))
op.set_dealias(True)
res["fused_dealias_ms"] = timeit(lambda: op.compute(v, ub, f, rho=rho, sens=sens))
for k in list(res):
    if k.endswith("_ms"): res[k.replace("_ms", "_gdofs")] = n / res[k] / 1e6
print(json.dumps(res))
