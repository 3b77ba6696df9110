"""Small end-to-end exercise of every kernel family for compute-sanitizer runs
(compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from helpers import Problem  # noqa: E402
from neko_top_b200 import operators as ops  # noqa: E402

for lx in (5, 6, 8):
    P = Problem(lx, ne=(3, 2, 2), deform=0.03)
    jacinv = (1.0 / P.t["jac"]).reshape(-1).cuda().contiguous()
    coef = ops.coef_t(ops.space_t(P.lx, P.D, P.w), P.nelv, P.cuda("G"), P.cuda("B"), jacinv)
    op = ops.fused_adjoint_rhs_t(coef)
    op.gs.init(P.keys.reshape(-1).cuda())
    v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
    f = [torch.zeros(P.n, device="cuda", dtype=torch.float64) for _ in range(3)]
    sens = torch.zeros(P.n, device="cuda", dtype=torch.float64)
    for mode in (0, 1):
        op.set_gs_mode(mode)
        op.step(v, ub, f, rho=rho, sens=sens)
    op.set_gs_mode(0)
    op.set_dealias(True)
    op.step(v, ub, f, rho=rho, sens=sens)
    op.set_dealias(False)
    adv = ops.adv_lin_b200_t(); adv.init(coef, op.handle)
    adv.compute_adjoint(*v, *ub, *f)
    adv.compute_linear(*v, *ub, *f)
    advd = ops.adv_lin_dealias_b200_t(); advd.init(None, coef, op.handle)
    advd.compute_adjoint(*v, *ub, *f)
    advd.compute_linear(*v, *ub, *f)
    mult = torch.ones(P.n, device="cuda", dtype=torch.float64)
    op.gs.op(mult)
    mult = 1.0 / mult
    w = [torch.zeros_like(sens) for _ in range(3)]
    Bs = P.cuda("B").clone(); op.gs.op(Bs)
    ops.curl(op, w, ub, jacinv, 1.0 / Bs)
    ops.min_dissipation_objective(op, *ub, rho, jacinv)
    flt = ops.PDE_filter_t(op, coef, mult, 0.1, abs_tol=1e-8, max_iter=50)
    flt.apply_forward(sens, rho)
    hv = [a.cpu().pin_memory() for a in v]; hub = [a.cpu().pin_memory() for a in ub]
    hf = [torch.empty(P.n, dtype=torch.float64).pin_memory() for _ in range(3)]
    op.step_host(hv, hub, rho.cpu().pin_memory(), hf, torch.empty(P.n, dtype=torch.float64).pin_memory())
    torch.cuda.synchronize()
    print("lx", lx, "ok", float(f[0].abs().sum()), flt.ksp_results[0])
    op.free()

# staged direct-stiffness summation at lx = 8 (levels 0/1/2): more elements than element slots so that the slots
# have runs with linked x faces, plus the y / z face passes and the set-up kernels; steady_simcomp update
P = Problem(8, ne=(8, 8, 8), deform=0.02)
coef = ops.coef_t(ops.space_t(P.lx, P.D, P.w), P.nelv, P.cuda("G"), P.cuda("B"))
op = ops.fused_adjoint_rhs_t(coef)
op.gs.init(P.keys.reshape(-1).cuda())
v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
f = [torch.zeros(P.n, device="cuda", dtype=torch.float64) for _ in range(3)]
for level in (0, 1, 2):
    op.set_xstage(level)
    op.step(v, ub, f, rho=rho)
    torch.cuda.synchronize()
    print("staged level", level, op.xstage_info(), float(f[0].abs().sum()))
sc = ops.steady_simcomp_t()
sc.init_from_attributes(1e-12, f)
sc.compute_()
print("steady", sc.normed_diff)
op.free()
