"""GPU parity of the fine-grid advection operators against the CPU
oracle: dealiased adjoint operator, linearised operator with and without dealiasing, the factory rule,
the fused right-hand side with dealiasing, and the discrete adjoint identity evaluated on the GPU.
fp64 fields: relative L2 <= 1e-12 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from helpers import Problem, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _ops():
    from neko_top_b200 import operators
    return operators


def _coef(P, with_jacinv=False):
    ops = _ops()
    jacinv = (1.0 / P.t["jac"]).reshape(-1).cuda().contiguous() if with_jacinv else None
    return ops.coef_t(ops.space_t(P.lx, P.D, P.w), P.nelv, P.cuda("G"), P.cuda("B"), jacinv)


def _case(lx, dealias):
    return {"case": {"numerics": {"dealias": dealias, "polynomial_order": lx - 1}}}


def _nan(n):
    return torch.full((n,), float("nan"), device="cuda", dtype=torch.float64)


@pytest.mark.parametrize("lx", [4, 5, 6, 7, 8, 9, 10])
def test_dealiased_adjoint_plugin(oracle, lx):
    """adv_lin_dealias_t%compute_adjoint (adv_adjoint_dealias.f90:235-462); f in/out; lxd = 3*lx/2."""
    ops = _ops()
    P = Problem(lx, ne=(2, 2, 3) if lx < 9 else (2, 2, 1), deform=0.03)
    lxd = 3 * lx // 2
    rng = np.random.default_rng(11)
    f0 = [rng.standard_normal(P.n) for _ in range(3)]
    fo = oracle.adjoint_advection_dealias(f0, P.v, P.ub, lx, lxd, P.nelv, P.G)
    adv = ops.advection_adjoint_factory(_case(lx, True), _coef(P))
    assert isinstance(adv, ops.adv_lin_dealias_b200_t)
    f = [torch.as_tensor(a).cuda() for a in f0]
    adv.compute_adjoint(*P.cuda("v"), *P.cuda("ub"), *f, n=P.n)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL, f"f[{c}] lx={lx}"
    adv.free()


@pytest.mark.parametrize("lx", [4, 5, 6, 7, 8, 9, 10])
def test_linear_plugin_no_dealias(oracle, lx):
    """adv_lin_no_dealias_t%compute_linear (adv_adjoint_no_dealias.f90:365-427); f in/out."""
    ops = _ops()
    P = Problem(lx, ne=(2, 3, 2), deform=0.03)
    rng = np.random.default_rng(12)
    f0 = [rng.standard_normal(P.n) for _ in range(3)]
    fo = oracle.linear_advection_no_dealias(f0, P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.jac)
    adv = ops.advection_adjoint_factory(_case(lx, False), _coef(P, with_jacinv=True))
    assert isinstance(adv, ops.adv_lin_b200_t)
    f = [torch.as_tensor(a).cuda() for a in f0]
    adv.compute_linear(*P.cuda("v"), *P.cuda("ub"), *f, n=P.n)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL, f"f[{c}] lx={lx}"
    adv.free()


@pytest.mark.parametrize("lx", [4, 5, 6, 7, 8, 9, 10])
def test_linear_plugin_dealias(oracle, lx):
    """adv_lin_dealias_t%compute_linear (adv_adjoint_dealias.f90:479-668); f in/out."""
    ops = _ops()
    P = Problem(lx, ne=(2, 2, 2) if lx < 9 else (2, 1, 1), deform=0.03)
    lxd = 3 * lx // 2
    rng = np.random.default_rng(13)
    f0 = [rng.standard_normal(P.n) for _ in range(3)]
    fo = oracle.linear_advection_dealias(f0, P.v, P.ub, lx, lxd, P.nelv, P.G)
    adv = ops.advection_adjoint_factory(_case(lx, True), _coef(P))
    f = [torch.as_tensor(a).cuda() for a in f0]
    adv.compute_linear(*P.cuda("v"), *P.cuda("ub"), *f, n=P.n)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL, f"f[{c}] lx={lx}"
    adv.free()


def test_factory_rule_and_errors():
    """advection_adjoint_fctry.f90:65-91: lxd from dealiased_polynomial_order, else 3*(order+1)/2; a
    fine order that is not instantiated fails loudly (no fallback)."""
    ops = _ops()
    from neko_top_b200 import _lib
    P = Problem(6, ne=(1, 1, 2), deform=0.0)
    coef = _coef(P)
    adv = ops.advection_adjoint_factory({"case": {"numerics": {"dealias": True, "polynomial_order": 5,
                                                               "dealiased_polynomial_order": 9}}}, coef)
    assert adv._hd.lxd == 9
    adv.free()
    _lib.set_abort_on_error(0)
    try:
        with pytest.raises(_lib.B200Error):
            ops.advection_adjoint_factory({"case": {"numerics": {"dealias": True, "polynomial_order": 5,
                                                                 "dealiased_polynomial_order": 11}}}, coef)
        op = ops.fused_adjoint_rhs_t(coef)
        with pytest.raises(_lib.B200Error):      # set_dealias before dealias_init
            _lib.check(_lib.lib().b200_adjrhs_set_dealias(op.handle.h, ops._ci(1)))
        op.free()
    finally:
        _lib.set_abort_on_error(1)


@pytest.mark.parametrize("lx", [6, 8])
def test_fused_rhs_dealias(oracle, lx):
    """case.numerics.dealias = true in the fused path: sources + mass matrix + DEALIASED adjoint advection
    + sensitivity (+ gs) against orc_adjoint_rhs with lxd > 0; BASELINE configs[2] runs this variant."""
    P = Problem(lx, ne=(3, 2, 2), deform=0.03)
    lxd = 3 * lx // 2
    ops = _ops()
    op = ops.fused_adjoint_rhs_t(_coef(P))
    op.set_dealias(True)
    v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
    fo, so, co = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho, lxd=lxd)
    f, sens, chi = [_nan(P.n) for _ in range(3)], _nan(P.n), _nan(P.n)
    op.compute(v, ub, f, rho=rho, sens=sens, chi_out=chi)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL
    assert rel_l2(sens.cpu().numpy(), so) <= TOL
    assert np.array_equal(chi.cpu().numpy(), co)
    # static forcing + given chi, then the whole step with gs
    rng = np.random.default_rng(3)
    chi_in = rng.random(P.n) * 1000.0
    fs = [rng.standard_normal(P.n) for _ in range(3)]
    fo, so, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, chi=chi_in, fstatic=fs, lxd=lxd)
    cid, nc = oracle.gs_classes(P.keys.reshape(-1).numpy())
    op.gs.init(P.keys.reshape(-1).cuda())
    f, sens = [_nan(P.n) for _ in range(3)], _nan(P.n)
    op.step(v, ub, f, chi=torch.as_tensor(chi_in).cuda(), fstatic=[torch.as_tensor(a).cuda() for a in fs], sens=sens)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), oracle.gs_add(fo[c], cid, nc)) <= TOL
    assert rel_l2(sens.cpu().numpy(), so) <= TOL
    # switching back gives the GLL-grid operator again
    op.set_dealias(False)
    fo, _, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho)
    f = [_nan(P.n) for _ in range(3)]
    op.compute(v, ub, f, rho=rho)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL
    op.free()


@pytest.mark.parametrize("dealias", [False, True])
def test_discrete_adjoint_identity_gpu(dealias):
    """<w, L v> == <A w, v> with L = compute_linear and A = compute_adjoint evaluated by the CUDA kernels on
    a deformed (non-affine) mesh: |diff| / (|w| |L v|) <= 1e-12 (SURVEY.md 8d correctness gates)."""
    ops = _ops()
    lx = 7
    P = Problem(lx, ne=(2, 2, 2), deform=0.05)
    Q = Problem(lx, ne=(2, 2, 2), deform=0.05, seed_shift=777)
    adv = ops.advection_adjoint_factory(_case(lx, dealias), _coef(P, with_jacinv=True))
    ub = P.cuda("ub")
    v, w = P.cuda("v"), Q.cuda("v")
    z = lambda: [torch.zeros(P.n, device="cuda", dtype=torch.float64) for _ in range(3)]
    Lv, Aw = z(), z()
    adv.compute_linear(*v, *ub, *Lv, n=P.n)
    adv.compute_adjoint(*w, *ub, *Aw, n=P.n)
    lhs = sum((a * b).sum().item() for a, b in zip(w, Lv))
    rhs = sum((a * b).sum().item() for a, b in zip(Aw, v))
    nw = np.sqrt(sum((a * a).sum().item() for a in w))
    nl = np.sqrt(sum((a * a).sum().item() for a in Lv))
    assert nl > 0 and abs(lhs - rhs) / (nw * nl) <= 1e-12
    adv.free()


@pytest.mark.parametrize("order", [2, 3])
def test_time_scheme_kernels(oracle, order):
    """sumab / makeabf / makebdf (adjoint_pnpn.f90:665-666,688-696) against the oracle's
    restatement of Neko's rhs_maker, and the one-pass makeabf+makebdf against the two calls back to back."""
    ops = _ops()
    rng = np.random.default_rng(21 + order)
    n = 3 * 7 ** 3 + 1            # odd length: exercises the tails
    r3 = lambda: [rng.standard_normal(n) for _ in range(3)]
    cu = lambda a: [torch.as_tensor(x).cuda() for x in a]
    u, l1, l2, f0, a1, a2 = r3(), r3(), r3(), r3(), r3(), r3()
    B = rng.random(n) + 0.5
    rho, dt = 1.3, 0.0125
    ab = [3.0, -3.0, 1.0] if order == 3 else [2.0, -1.0, 0.0]
    bd = [11.0 / 6.0, 3.0, -1.5, 1.0 / 3.0] if order == 3 else [1.5, 2.0, -0.5, 0.0]
    # sumab
    ref = oracle.sumab(u, l1, l2, ab, order)
    ue = [torch.full((n,), float("nan"), device="cuda", dtype=torch.float64) for _ in range(3)]
    du, dl1, dl2 = cu(u), cu(l1), cu(l2)
    lag = lambda c: (dl1[c], dl2[c])
    ops.rhs_maker_sumab_t().compute_fluid(*ue, *du, lag(0), lag(1), lag(2), ab, order)
    for c in range(3):
        assert rel_l2(ue[c].cpu().numpy(), ref[c]) <= TOL
    # makeabf then makebdf
    ra1, ra2, rf = oracle.makeabf(a1, a2, f0, rho, ab)
    rf2 = oracle.makebdf(l1, l2, rf, u, B, rho, dt, bd, order)
    df, da1, da2, dB = cu(f0), cu(a1), cu(a2), torch.as_tensor(B).cuda()
    ops.rhs_maker_ext_t().compute_fluid(*da1, *da2, *df, rho, ab)
    for c in range(3):
        assert rel_l2(df[c].cpu().numpy(), rf[c]) <= TOL
        assert np.array_equal(da1[c].cpu().numpy(), f0[c]) and np.array_equal(da2[c].cpu().numpy(), a1[c])
    ops.rhs_maker_bdf_t().compute_fluid(lag(0), lag(1), lag(2), *df, *du, dB, rho, dt, bd, order)
    for c in range(3):
        assert rel_l2(df[c].cpu().numpy(), rf2[c]) <= TOL
    # one pass
    gf, ga1, ga2 = cu(f0), cu(a1), cu(a2)
    ops.makeabf_bdf(ga1, ga2, lag(0), lag(1), lag(2), gf, du, dB, rho, dt, ab, bd, order)
    for c in range(3):      # (FMA contraction may differ between one and two passes: last-bit differences in f)
        assert rel_l2(gf[c].cpu().numpy(), rf2[c]) <= TOL
        assert torch.equal(ga1[c], da1[c]) and torch.equal(ga2[c], da2[c])


@pytest.mark.parametrize("lx", [5, 8])
def test_min_dissipation_chain(oracle, lx):
    """Neko curl (strong curl, B, gs, Binv), the curl-curl forcing of
    adjoint_minimum_dissipation_source_term.f90:231-243 with and without a point-zone mask, the objective of
    minimum_dissipation_objective_function.f90:186-254 and mask_exterior_const on the device."""
    ops = _ops()
    P = Problem(lx, ne=(3, 2, 2), deform=0.03)
    keys = P.keys.reshape(-1).numpy()
    cid, nc = oracle.gs_classes(keys)
    Bsum = oracle.gs_add(P.B, cid, nc)
    Binv, jacinv = 1.0 / Bsum, 1.0 / P.jac
    coef = _coef(P, with_jacinv=True)
    op = ops.fused_adjoint_rhs_t(coef)
    op.gs.init(P.keys.reshape(-1).cuda())
    dBinv = torch.as_tensor(Binv).cuda()
    u = P.cuda("ub")
    # curl
    wref = oracle.curl(P.ub, lx, P.nelv, P.D, P.G, jacinv, P.B, Binv, cid, nc)
    w = [_nan(P.n) for _ in range(3)]
    ops.curl(op, w, u, coef.jacinv, dBinv)
    for c in range(3):
        assert rel_l2(w[c].cpu().numpy(), wref[c]) <= TOL
    # curl-curl forcing, unmasked and masked
    rng = np.random.default_rng(9)
    f0 = [rng.standard_normal(P.n) for _ in range(3)]
    mask = (np.sort(rng.choice(P.n, P.n // 5, replace=False)) + 1).astype(np.int32)
    dmask = torch.as_tensor(mask).cuda()
    for mk, dm in ((None, None), (mask, dmask)):
        ref = oracle.curlcurl_forcing(f0, P.ub, lx, P.nelv, P.D, P.G, jacinv, P.B, Binv, cid, nc, mask=mk, obj_scale=0.7)
        f = [torch.as_tensor(a).cuda() for a in f0]
        st = ops.adjoint_minimum_dissipation_source_term_t()
        st.init_from_components(*f, *u, 0.7, dm, dm is not None, coef, op, dBinv)
        st.compute_()
        for c in range(3):
            assert rel_l2(f[c].cpu().numpy(), ref[c]) <= TOL
    # objective
    chi = rng.random(P.n) * 1000.0
    for mk, dm in ((None, None), (mask, dmask)):
        ref = oracle.min_dissipation_objective(P.ub, chi, lx, P.nelv, P.D, P.G, jacinv, P.B, mask=mk, K=2.0, obj_scale=0.5)
        got = ops.min_dissipation_objective(op, *u, torch.as_tensor(chi).cuda(), coef.jacinv, mask=dm, K=2.0, obj_scale=0.5)
        for a, b in zip(got, ref):
            assert abs(a - b) <= 1e-12 * max(1.0, abs(b))
    # mask_exterior_const
    fld = torch.as_tensor(f0[0]).cuda()
    ops.mask_exterior_const(fld, dmask, -3.5)
    assert np.array_equal(fld.cpu().numpy(), oracle.mask_exterior_const(f0[0], mask, -3.5))
    op.free()


@pytest.mark.parametrize("lx,precond", [(5, "jacobi"), (5, "ident"), (8, "jacobi")])
def test_pde_filter(oracle, lx, precond):
    """PDE_filter_t%apply (PDE_filter_mapping.f90:212-282): CG on (r^2 K + M) x = gs(B x_in)
    with ax_helm + gs, against a dense LAPACK solve of the operator assembled from the oracle's ax_helm on a
    small deformed mesh; a constant field is a fixed point of the filter; forward and backward use the same
    operator.  Converged solves (residual 1e-14) agree with the dense solve to 1e-12 (measured 1.5e-13 .. 3.4e-13);
    starting from the unfiltered field (PDE_filter_mapping.f90:246-248) gives the same answer in no more
    iterations; the reference's own settings (abstol 1e-10, <= 200 iterations, :131-137) reach 1e-7."""
    ops = _ops()
    P = Problem(lx, ne=(2, 2, 2) if lx < 8 else (2, 1, 1), deform=0.03)
    keys = P.keys.reshape(-1).numpy()
    cid, nc = oracle.gs_classes(keys)
    jacinv = 1.0 / P.jac
    mult = 1.0 / oracle.gs_add(np.ones(P.n), cid, nc)
    coef = _coef(P, with_jacinv=True)
    op = ops.fused_adjoint_rhs_t(coef)
    op.gs.init(P.keys.reshape(-1).cuda())
    radius = 0.08
    flt = ops.PDE_filter_t(op, coef, torch.as_tensor(mult).cuda(), radius, abs_tol=1e-14, max_iter=2000, precond=precond)
    ref = oracle.pde_filter_dense(P.rho, lx, P.nelv, P.D, P.w, P.G, jacinv, P.B, cid, nc, radius)
    x_in, x_out = P.cuda("rho"), _nan(P.n)
    flt.apply_forward(x_out, x_in)
    iters, r0, r1 = flt.ksp_results
    assert 0 < iters < 2000 and r1 < 1e-14 <= r0
    err0 = rel_l2(x_out.cpu().numpy(), ref)
    assert err0 <= 1e-12, err0
    # same solve started from the unfiltered field: same answer, not more iterations; bit-reproducible
    flt.x0_is_input = True
    x2 = _nan(P.n)
    flt.apply_forward(x2, x_in)
    iters2, _, r12 = flt.ksp_results
    assert 0 < iters2 <= iters and r12 < 1e-14
    assert rel_l2(x2.cpu().numpy(), ref) <= 1e-12
    x3 = _nan(P.n)
    flt.apply_forward(x3, x_in)
    assert torch.equal(x2, x3) and flt.ksp_results[0] == iters2
    flt.x0_is_input = False
    one = torch.ones(P.n, device="cuda", dtype=torch.float64)
    flt.apply_forward(x_out, one)
    assert float((x_out - 1.0).abs().max()) <= 1e-11
    g = _nan(P.n)
    flt.apply_backward(g, x_in)
    assert rel_l2(g.cpu().numpy(), ref) <= 1e-12
    # the reference's settings: abstol 1e-10, at most 200 iterations, identity preconditioner
    ref_flt = ops.PDE_filter_t(op, coef, torch.as_tensor(mult).cuda(), radius, x0_is_input=True)
    ref_flt.apply_forward(x_out, x_in)
    it_r, _, r_r = ref_flt.ksp_results
    assert it_r <= 200 and (r_r < 1e-10 or it_r == 200)
    if r_r < 1e-10:
        assert rel_l2(x_out.cpu().numpy(), ref) <= 1e-7
    print(f"pde_filter lx={lx} {precond}: iters {iters} (x0=0) / {iters2} (x0=x_in), err vs dense {err0:.2e}; "
          f"reference settings: {it_r} iterations, residual {r_r:.2e}")
    op.free()


def test_steady_simcomp(oracle):
    """steady_simcomp_t%compute_ (simulation_components/steady_simcomp.f90:154-188) through
    b200_steady_field_update: squared norm of the change per field against the oracle restatement, old <- new,
    freeze once below the tolerance, and a run-to-run deterministic reduction (same bits)."""
    from oracle import np_oracle as npo
    ops = _ops()
    rng = np.random.default_rng(21)
    n = 3 * 8 ** 3 * 37 + 5                      # not a multiple of anything in the kernel
    new = [rng.standard_normal(n) for _ in range(4)]          # u, v, w, p
    old0 = [a + 1e-4 * rng.standard_normal(n) for a in new]
    d_new = [torch.as_tensor(a).cuda() for a in new]
    sc = ops.steady_simcomp_t()
    sc.init_from_attributes(1e-12, d_new)
    for o, src in zip(sc.old, old0):
        o.copy_(torch.as_tensor(src))
    old_np = [a.copy() for a in old0]
    nd_ref, fr_ref = npo.steady_simcomp_compute(new, old_np, 1e-12)
    sc.compute_()
    assert sc.freeze == fr_ref == False
    for a, b in zip(sc.normed_diff, nd_ref):
        assert abs(a - b) <= 1e-13 * abs(b)
    for o, a in zip(sc.old, d_new):
        assert torch.equal(o, a)
    # determinism: the same update from the same state gives the same bits
    vals = []
    for _ in range(3):
        for o, src in zip(sc.old, old0):
            o.copy_(torch.as_tensor(src))
        sc.compute_()
        vals.append(tuple(sc.normed_diff))
    assert vals[0] == vals[1] == vals[2]
    # unchanged fields -> zero change -> freeze, and a frozen component does nothing
    sc.compute_()
    assert sc.freeze and max(sc.normed_diff) == 0.0
    # empty field
    e = torch.empty(0, device="cuda", dtype=torch.float64)
    sc2 = ops.steady_simcomp_t()
    sc2.init_from_attributes(1.0, [e])
    sc2.compute_()
    assert sc2.normed_diff == [0.0] and sc2.freeze


def test_step_host_dealias_chunks(oracle):
    """b200_adjrhs_step_host with the dealiased operator and several element chunks (nelv >= 512): every chunk
    must compute ITS elements (the fine-grid operator honours the chunk's first element), result identical to
    the device-resident step and within 1e-12 of the oracle."""
    lx = 6
    P = Problem(lx, ne=(8, 8, 9), deform=0.02)            # 576 elements -> 2 chunks
    ops = _ops()
    coef = ops.coef_t(ops.space_t(P.lx, P.D, P.w), P.nelv, P.cuda("G"), P.cuda("B"))
    op = ops.fused_adjoint_rhs_t(coef)
    op.gs.init(P.keys.reshape(-1).cuda())
    op.set_dealias(True)
    v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
    nan = lambda: torch.full((P.n,), float("nan"), device="cuda", dtype=torch.float64)
    f, sens = [nan() for _ in range(3)], nan()
    op.step(v, ub, f, rho=rho, sens=sens)
    hv = [a.cpu().pin_memory() for a in v]
    hub = [a.cpu().pin_memory() for a in ub]
    hf = [torch.full((P.n,), float("nan"), dtype=torch.float64).pin_memory() for _ in range(3)]
    hs = torch.full((P.n,), float("nan"), dtype=torch.float64).pin_memory()
    op.step_host(hv, hub, rho.cpu().pin_memory(), hf, hs)
    for c in range(3):
        assert torch.equal(hf[c], f[c].cpu()), "host-buffer step (chunked, dealiased) differs from the device step"
    assert torch.equal(hs, sens.cpu())
    fo, so, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho, lxd=3 * lx // 2)
    cid, nc = oracle.gs_classes(P.keys.reshape(-1).numpy())
    for c in range(3):
        assert rel_l2(hf[c].numpy(), oracle.gs_add(fo[c], cid, nc)) <= 1e-12
    assert rel_l2(hs.numpy(), so) <= 1e-12
    op.free()


def test_dealias_lx8_tensor_core_kernel_many_elements(oracle):
    """lx = 8 / lxd = 12: the dealiased adjoint operator on the FP64 tensor cores (advop_mma_kernel) with more elements
    than SMs (every CTA loops, the next element's fields are prefetched), in mesh order, in a permuted element order
    (element list) and through the chunked host-buffer step; the un-fused accumulate drop-in on the same mesh.
    All <= 1e-12 against the oracle's dealiased operator."""
    lx, lxd = 8, 12
    P = Problem(lx, ne=(8, 8, 9), deform=0.03)            # 576 elements: ~4 per CTA, 2 chunks in step_host
    ops = _ops()
    coef = ops.coef_t(ops.space_t(P.lx, P.D, P.w), P.nelv, P.cuda("G"), P.cuda("B"))
    op = ops.fused_adjoint_rhs_t(coef)
    op.gs.init(P.keys.reshape(-1).cuda())
    op.set_dealias(True)
    v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
    nan = lambda: torch.full((P.n,), float("nan"), device="cuda", dtype=torch.float64)
    fo, so, co = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho, lxd=lxd)
    f, sens, chi = [nan() for _ in range(3)], nan(), nan()
    op.compute(v, ub, f, rho=rho, sens=sens, chi_out=chi)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= 1e-12
    assert rel_l2(sens.cpu().numpy(), so) <= 1e-12
    assert np.array_equal(chi.cpu().numpy(), co)
    # permuted processing order: same bits (the elements are independent)
    op.set_element_order(np.random.default_rng(11).permutation(P.nelv))
    g = [nan() for _ in range(3)]
    op.compute(v, ub, g, rho=rho)
    for c in range(3):
        assert torch.equal(f[c], g[c])
    op.set_element_order(None)
    # step (+ summation) and the chunked host-buffer step
    cid, nc = oracle.gs_classes(P.keys.reshape(-1).numpy())
    op.step(v, ub, f, rho=rho, sens=sens)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), oracle.gs_add(fo[c], cid, nc)) <= 1e-12
    hv = [a.cpu().pin_memory() for a in v]
    hub = [a.cpu().pin_memory() for a in ub]
    hf = [torch.full((P.n,), float("nan"), dtype=torch.float64).pin_memory() for _ in range(3)]
    hs = torch.full((P.n,), float("nan"), dtype=torch.float64).pin_memory()
    op.step_host(hv, hub, rho.cpu().pin_memory(), hf, hs)
    op.set_xstage(0)
    op.step(v, ub, g, rho=rho)
    for c in range(3):
        assert torch.equal(hf[c], g[c].cpu()), "host-buffer step (chunked, dealiased) differs from the device step"
    # un-fused drop-in: f is in/out
    adv = ops.adv_lin_dealias_b200_t(); adv.init(None, coef, op.handle)
    rng = np.random.default_rng(8)
    f0 = [rng.standard_normal(P.n) for _ in range(3)]
    fa = [torch.as_tensor(a).cuda() for a in f0]
    adv.compute_adjoint(*v, *ub, *fa)
    fr = oracle.adjoint_advection_dealias(f0, P.v, P.ub, lx, lxd, P.nelv, P.G)
    for c in range(3):
        assert rel_l2(fa[c].cpu().numpy(), fr[c]) <= 1e-12
    op.free()


def test_bench_dealiased_leg():
    """bench.py's extra `dealiased` entry: times the dealiased step and checks the fused right-hand side of a sample
    against the oracle."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    d = bench.dealiased_leg(6, 8, torch.device("cuda", 0), steps=2, sample_nel=32)
    assert d["parity"]["ok"] and d["value"] > 0 and "advop_mma_kernel" in d["kernel"]
