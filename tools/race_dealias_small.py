"""Fused dealiased step at lx = 8 on 4 elements, for `compute-sanitizer --tool racecheck` (advop_mma_kernel's in-place
shared-memory GEMM tiles)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import Problem
from neko_top_b200 import operators as ops
P = Problem(8, ne=(2, 2, 1), deform=0.03)
coef = ops.coef_t(ops.space_t(P.lx, P.D, P.w), P.nelv, P.cuda("G"), P.cuda("B"))
op = ops.fused_adjoint_rhs_t(coef)
v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
f = [torch.zeros(P.n, device="cuda", dtype=torch.float64) for _ in range(3)]
op.set_dealias(True)
op.compute(v, ub, f, rho=rho)
torch.cuda.synchronize()
print("ok")
