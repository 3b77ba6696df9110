"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  fp64 fields: relative L2 <= 1e-12 (BASELINE.json north_star); gather-scatter index
maps: exact integer equality."""
import numpy as np
import pytest
import torch

from helpers import Problem, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _ops():
    from neko_top_b200 import operators
    return operators


def _fused(P):
    ops = _ops()
    coef = ops.coef_t(ops.space_t(P.lx, P.D, P.w), P.nelv, P.cuda("G"), P.cuda("B"))
    return ops.fused_adjoint_rhs_t(coef), coef


@pytest.mark.parametrize("lx", [4, 5, 6, 7, 8, 9, 10])
def test_fused_rhs_matches_oracle(oracle, lx):
    P = Problem(lx, ne=(3, 2, 2), deform=0.03)
    fo, so, co = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho)
    op, _ = _fused(P)
    v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
    f = [torch.full((P.n,), float("nan"), device="cuda", dtype=torch.float64) for _ in range(3)]
    sens = torch.full((P.n,), float("nan"), device="cuda", dtype=torch.float64)
    chi = torch.full((P.n,), float("nan"), device="cuda", dtype=torch.float64)
    op.compute(v, ub, f, rho=rho, sens=sens, chi_out=chi)
    torch.cuda.synchronize()
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL, f"f[{c}] lx={lx}"
    assert rel_l2(sens.cpu().numpy(), so) <= TOL
    assert np.array_equal(chi.cpu().numpy(), co), "RAMP map must be bit-exact"
    op.free()
