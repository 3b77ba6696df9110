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


def _nan(n):
    return torch.full((n,), float("nan"), device="cuda", dtype=torch.float64)


@pytest.mark.parametrize("lx", [6, 8])
def test_fused_variants(oracle, lx):
    """chi given instead of rho; static forcing; lube off; convex-down RAMP; masked lube."""
    P = Problem(lx, ne=(2, 3, 2), deform=0.03)
    rng = np.random.default_rng(5)
    op, _ = _fused(P)
    v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
    # (a) chi_in + f_static
    chi = rng.random(P.n) * 1000.0
    fs = [rng.standard_normal(P.n) for _ in range(3)]
    fo, so, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, chi=chi, fstatic=fs)
    f, sens = [_nan(P.n) for _ in range(3)], _nan(P.n)
    op.compute(v, ub, f, chi=torch.as_tensor(chi).cuda(), fstatic=[torch.as_tensor(a).cuda() for a in fs], sens=sens)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL
    assert rel_l2(sens.cpu().numpy(), so) <= TOL
    # (b) lube off, convex-down RAMP, other constants
    op.set_params(f_min=2.0, f_max=500.0, q=0.5, convex_up=False, if_lube=False)
    fo, so, co = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho, f_min=2.0,
                                    f_max=500.0, q=0.5, convex_up=0, if_lube=0)
    f, sens, chio = [_nan(P.n) for _ in range(3)], _nan(P.n), _nan(P.n)
    op.compute(v, ub, f, rho=rho, sens=sens, chi_out=chio)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL
    assert rel_l2(sens.cpu().numpy(), so) <= TOL
    assert np.array_equal(chio.cpu().numpy(), co)
    # (c) masked lube (1-based point-zone indices), K = 2.5
    op.set_params(if_lube=True, K_lube=2.5, K_sens=2.5)
    mask = (np.sort(rng.choice(P.n, P.n // 7, replace=False)) + 1).astype(np.int32)
    op.set_lube_mask(torch.as_tensor(mask).cuda())
    fo, so, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho, mask=mask,
                                   K_lube=2.5, K_sens=2.5)
    f, sens = [_nan(P.n) for _ in range(3)], _nan(P.n)
    op.compute(v, ub, f, rho=rho, sens=sens)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL
    assert rel_l2(sens.cpu().numpy(), so) <= TOL
    op.free()


@pytest.mark.parametrize("lx", [5, 8])
def test_advection_adjoint_plugin(oracle, lx):
    """advection_adjoint_t%compute_adjoint drop-in: f is in/out (accumulated)."""
    ops = _ops()
    P = Problem(lx, ne=(2, 2, 2), deform=0.03)
    rng = np.random.default_rng(8)
    f0 = [rng.standard_normal(P.n) for _ in range(3)]
    fo = oracle.adjoint_advection_no_dealias(f0, P.v, P.ub, lx, P.nelv, P.D, P.w, P.G)
    coef = ops.coef_t(ops.space_t(P.lx, P.D, P.w), P.nelv, P.cuda("G"), P.cuda("B"))
    adv = ops.advection_adjoint_factory({"case": {"numerics": {"dealias": False, "polynomial_order": lx - 1}}}, coef)
    f = [torch.as_tensor(a).cuda() for a in f0]
    v, ub = P.cuda("v"), P.cuda("ub")
    adv.compute_adjoint(*v, *ub, *f, n=P.n)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL
    adv.free()


def test_unfused_pointwise_plugins(oracle):
    """The reference's call order with the un-fused drop-ins: source_term%compute (Brinkman, lube),
    opcolv, compute_adjoint, compute_sensitivity == the fused kernel == the oracle."""
    ops = _ops()
    lx = 7
    P = Problem(lx, ne=(2, 2, 3), deform=0.03)
    fo, so, co = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho)
    v, ub, rho, B = P.cuda("v"), P.cuda("ub"), P.cuda("rho"), P.cuda("B")
    chi = _nan(P.n)
    ops.RAMP_mapping_t().apply_forward(chi, rho)
    assert np.array_equal(chi.cpu().numpy(), co)
    f = [torch.zeros(P.n, device="cuda", dtype=torch.float64) for _ in range(3)]
    br = ops.simple_brinkman_source_term_t()
    br.init_from_components(*f, chi, *v, None)
    br.compute_()
    lu = ops.adjoint_lube_source_term_t()
    lu.init_from_components(*f, chi, 1.0, *ub)
    lu.compute_()
    ops.opcolv(*f, B)
    coef = ops.coef_t(ops.space_t(P.lx, P.D, P.w), P.nelv, P.cuda("G"), B)
    adv = ops.advection_adjoint_factory({"case": {"numerics": {"dealias": False}}}, coef)
    adv.compute_adjoint(*v, *ub, *f)
    sens = _nan(P.n)
    ops.compute_sensitivity(sens, *ub, *v)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL
    assert rel_l2(sens.cpu().numpy(), so) <= TOL
    # RAMP backward
    dF = torch.as_tensor(np.linspace(-1, 1, P.n)).cuda()
    out = _nan(P.n)
    ops.RAMP_mapping_t().apply_backward(out, dF, rho)
    assert rel_l2(out.cpu().numpy(), oracle.ramp_backward(dF.cpu().numpy(), P.rho)) <= 1e-15
    adv.free()


@pytest.mark.parametrize("lx,ne", [(4, (3, 2, 2)), (8, (3, 3, 2)), (5, (1, 1, 1))])
def test_gs_index_maps_bit_exact(oracle, lx, ne):
    """Node classes of the GPU gather-scatter == the oracle's, after the canonical relabelling
    (exact integer comparison)."""
    P = Problem(lx, ne=ne, deform=0.0)
    op, _ = _fused(P)
    key = P.keys.reshape(-1)
    op.gs.init(key.cuda())
    cid, nc = op.gs.classes()
    cido, nco = oracle.gs_classes(key.numpy())
    assert nc == nco and np.array_equal(cid, cido)
    # host-key path gives the same map
    op.gs.init(key.numpy())
    cid2, nc2 = op.gs.classes()
    assert nc2 == nco and np.array_equal(cid2, cido)
    op.free()


def test_gs_op_parity_and_properties(oracle):
    lx = 6
    P = Problem(lx, ne=(3, 2, 2), deform=0.0)
    op, _ = _fused(P)
    key = P.keys.reshape(-1)
    op.gs.init(key.cuda())
    cido, nco = oracle.gs_classes(key.numpy())
    rng = np.random.default_rng(1)
    f = [rng.standard_normal(P.n) for _ in range(3)]
    g = [torch.as_tensor(a).cuda() for a in f]
    op.gs.op3(*g)
    for c in range(3):
        assert np.array_equal(g[c].cpu().numpy(), oracle.gs_add(f[c], cido, nco)), "same summation order => bit-exact"
    one = torch.as_tensor(f[0]).cuda()
    op.gs.op(one)
    assert np.array_equal(one.cpu().numpy(), g[0].cpu().numpy())
    mult = torch.ones(P.n, device="cuda", dtype=torch.float64)
    op.gs.op(mult)
    assert mult.min().item() == 1 and mult.max().item() == 8
    again = (g[1] / mult).clone()
    op.gs.op(again)
    assert rel_l2(again.cpu().numpy(), g[1].cpu().numpy()) <= 1e-14
    op.free()


@pytest.mark.parametrize("lx", [6, 8])
def test_step_matches_oracle(oracle, lx):
    """One bench "step": fused kernel + gs_op on f."""
    P = Problem(lx, ne=(3, 3, 2), deform=0.03)
    fo, so, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho)
    cid, nc = oracle.gs_classes(P.keys.reshape(-1).numpy())
    op, _ = _fused(P)
    op.gs.init(P.keys.reshape(-1).cuda())
    op.set_xstage(1)       # levels 0 and 1 are bit-identical to the host-buffer path; level 2: test_step_xstage
    v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
    f, sens = [_nan(P.n) for _ in range(3)], _nan(P.n)
    op.step(v, ub, f, rho=rho, sens=sens)
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), oracle.gs_add(fo[c], cid, nc)) <= TOL
    # host-buffer entry point (bench e2e) gives the identical result
    hv = [a.cpu().pin_memory() for a in v]
    hub = [a.cpu().pin_memory() for a in ub]
    hf = [torch.empty(P.n, dtype=torch.float64).pin_memory() for _ in range(3)]
    hs = torch.empty(P.n, dtype=torch.float64).pin_memory()
    op.step_host(hv, hub, rho.cpu().pin_memory(), hf, hs)
    for c in range(3):
        assert torch.equal(hf[c], f[c].cpu())
    assert torch.equal(hs, sens.cpu())
    op.free()


def test_golden_fixture_gpu():
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adjrhs_lx4.json")))
    P = Problem(g["lx"], ne=tuple(g["ne"]), deform=g["deform"])
    op, _ = _fused(P)
    f, sens, chi = [_nan(P.n) for _ in range(3)], _nan(P.n), _nan(P.n)
    op.compute(P.cuda("v"), P.cuda("ub"), f, rho=P.cuda("rho"), sens=sens, chi_out=chi)
    idx = np.asarray(g["idx"])
    for c in range(3):
        assert np.allclose(f[c].cpu().numpy()[idx], g["f"][c], rtol=1e-12, atol=1e-13)
    assert np.allclose(sens.cpu().numpy()[idx], g["sens"], rtol=1e-13, atol=1e-14)
    assert np.allclose(chi.cpu().numpy()[idx], g["chi"], rtol=1e-14)
    op.free()


def test_duct_configs(oracle):
    """BASELINE configs[0] (lx=6) and configs[2] (lx=8): duct 24x8x8, chi = 1000 in the lowperm box."""
    from neko_top_b200 import sem, workloads
    ops = _ops()
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
        op = ops.fused_adjoint_rhs_t(ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), brick.nelv, Gf, Bf))
        op.gs.init(flat(keys))
        f, sens = [_nan(brick.n) for _ in range(3)], _nan(brick.n)
        op.step(v, ub, f, chi=chif, sens=sens)
        c = lambda t: t.cpu().numpy()
        fo, so, _ = oracle.adjoint_rhs([c(a) for a in v], [c(a) for a in ub], lx, brick.nelv, sp.dx, sp.wx,
                                       [c(g) for g in Gf], c(Bf), chi=c(chif))
        cid, nc = oracle.gs_classes(c(flat(keys)))
        for k in range(3):
            assert rel_l2(c(f[k]), oracle.gs_add(fo[k], cid, nc)) <= TOL
        assert rel_l2(c(sens), so) <= TOL
        if lx == 8:      # configs[2] ships `dealias: true` (laminar.case:9-13): lxd = 12
            op.set_dealias(True)
            op.step(v, ub, f, chi=chif, sens=sens)
            fo, so, _ = oracle.adjoint_rhs([c(a) for a in v], [c(a) for a in ub], lx, brick.nelv, sp.dx, sp.wx,
                                           [c(g) for g in Gf], c(Bf), chi=c(chif), lxd=12)
            for k in range(3):
                assert rel_l2(c(f[k]), oracle.gs_add(fo[k], cid, nc)) <= TOL
            assert rel_l2(c(sens), so) <= TOL
        op.free()


def test_full_size_properties():
    """BASELINE configs[1] (32^3, lx=8) through size-independent properties: linearity in the adjoint
    velocity, sum conservation of the direct-stiffness sum, determinism."""
    from neko_top_b200 import sem, workloads
    ops = _ops()
    lx, ne = 8, 32
    brick = workloads.config_box(ne, lx)
    sp = sem.Space(lx)
    x, y, z = workloads.coords(brick, "cuda")
    keys = workloads.node_keys(brick, "cuda")
    G, _, B = sem.geometric_factors(x, y, z, sp)
    fl = workloads.make_fields(brick, x, y, z, keys)
    fl2 = workloads.make_fields(brick, x, y, z, keys + 12345)
    del x, y, z
    flat = lambda a: a.reshape(-1).contiguous()
    op = ops.fused_adjoint_rhs_t(ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), brick.nelv, [flat(g) for g in G], flat(B)))
    op.gs.init(flat(keys))
    v1, v2, ub, rho = [flat(a) for a in fl.v], [flat(a) for a in fl2.v], [flat(a) for a in fl.ub], flat(fl.rho)
    n = brick.n
    mk = lambda: [torch.empty(n, device="cuda", dtype=torch.float64) for _ in range(3)]
    a, b = 0.75, -1.5
    f1, f2, fm = mk(), mk(), mk()
    op.set_params(if_lube=False)       # the lube term K*chi*u_b does not depend on the adjoint velocity
    op.compute(v1, ub, f1, rho=rho)
    op.compute(v2, ub, f2, rho=rho)
    op.compute([a * p + b * q for p, q in zip(v1, v2)], ub, fm, rho=rho)
    for c in range(3):
        ref = a * f1[c] + b * f2[c]
        err = (torch.linalg.norm(fm[c] - ref) / torch.linalg.norm(ref)).item()
        assert err <= 1e-11, f"linearity violated: {err:.3e}"   # a*v1+b*v2 is itself rounded; see below
    # determinism
    f1b = mk()
    op.compute(v1, ub, f1b, rho=rho)
    assert all(torch.equal(p, q) for p, q in zip(f1, f1b))
    # gs conserves the multiplicity-weighted sum and makes the field C0
    mult = torch.ones(n, device="cuda", dtype=torch.float64)
    op.gs.op(mult)
    s0 = f1[0].sum().item()
    g = f1[0].clone()
    op.gs.op(g)
    assert abs((g / mult).sum().item() - s0) <= 1e-9 * f1[0].abs().sum().item()
    k = flat(keys)
    order = torch.argsort(k)
    ks, gsrt = k[order], g[order]
    same = ks[1:] == ks[:-1]
    assert torch.equal(gsrt[1:][same], gsrt[:-1][same])
    op.free()


def test_config1_full_size_vs_oracle(oracle):
    """BASELINE configs[1] at FULL size (32^3 elements, lx = 8, 16.8 MDOF; random design field, RAMP 0/1000/1
    convex-up, K = 1) compared directly with the oracle: f after gs_op and the sensitivity <= 1e-12, and the
    gather-scatter class map integer-identical on the whole mesh."""
    from neko_top_b200 import sem, workloads
    ops = _ops()
    lx, ne = 8, 32
    brick = workloads.config_box(ne, lx)
    sp = sem.Space(lx)
    x, y, z = workloads.coords(brick, "cuda")
    keys = workloads.node_keys(brick, "cuda")
    G, _, B = sem.geometric_factors(x, y, z, sp)
    fl = workloads.make_fields(brick, x, y, z, keys)
    del x, y, z
    flat = lambda a: a.reshape(-1).contiguous()
    Gf, Bf = [flat(g) for g in G], flat(B)
    v, ub, rho, kf = [flat(a) for a in fl.v], [flat(a) for a in fl.ub], flat(fl.rho), flat(keys)
    op = ops.fused_adjoint_rhs_t(ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), brick.nelv, Gf, Bf))
    op.set_params()
    op.gs.init(kf)
    n = brick.n
    f, sens = [_nan(n) for _ in range(3)], _nan(n)
    op.step(v, ub, f, rho=rho, sens=sens)
    active, nstaged, nleft, ntot = op.xstage_info()
    assert active == 2 and nleft <= 0.12 * ntot
    c = lambda t: t.cpu().numpy()
    fo, so, _ = oracle.adjoint_rhs([c(a) for a in v], [c(a) for a in ub], lx, brick.nelv, sp.dx, sp.wx,
                                   [c(g) for g in Gf], c(Bf), rho=c(rho))
    cid, nc = oracle.gs_classes(c(kf))
    gcid, gnc = op.gs.classes()
    assert gnc == nc and np.array_equal(gcid, cid), "gather-scatter index map differs from the oracle's at 32^3"
    for k in range(3):
        assert rel_l2(c(f[k]), oracle.gs_add(fo[k], cid, nc)) <= TOL
    assert rel_l2(c(sens), so) <= TOL
    op.free()


def _check_step_host(op, v, ub, rho, f, sens):
    """b200_adjrhs_step_host (host buffers; several element chunks, summation and copies pipelined when the
    elements are in mesh order) must reproduce the device-resident step bit for bit."""
    n = f[0].numel()
    hv = [a.cpu().pin_memory() for a in v]
    hub = [a.cpu().pin_memory() for a in ub]
    hf = [torch.full((n,), float("nan"), dtype=torch.float64).pin_memory() for _ in range(3)]
    hs = torch.full((n,), float("nan"), dtype=torch.float64).pin_memory() if sens is not None else None
    for _ in range(2):
        op.step_host(hv, hub, rho.cpu().pin_memory(), hf, hs)
    for c in range(3):
        assert torch.equal(hf[c], f[c].cpu()), "host-buffer step differs from the device step"
    if sens is not None:
        assert torch.equal(hs, sens.cpu())


@pytest.mark.parametrize("order_kind", ["mesh", "tile", "random"])
def test_step_gs_packed_lists(oracle, order_kind):
    """lx = 8: b200_adjrhs_step with the class lists packed by size and sorted by completing element (gs mode 1: the
    lists the pipelined host step walks), any processing order.  Against the oracle <= 1e-12 and BIT-identical to
    the CSR pass; repeated steps reuse the schedule."""
    from neko_top_b200 import workloads
    lx = 8
    P = Problem(lx, ne=(12, 10, 9), deform=0.02)      # 1080 elements = 3 windows of 444 slots
    fo, so, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho)
    cid, nc = oracle.gs_classes(P.keys.reshape(-1).numpy())
    ref = [oracle.gs_add(fo[c], cid, nc) for c in range(3)]
    op, _ = _fused(P)
    op.gs.init(P.keys.reshape(-1).cuda())
    op.set_xstage(1)
    if order_kind == "tile":
        op.set_element_order(workloads.tile_order(P.brick, (4, 4)))
    elif order_kind == "random":
        op.set_element_order(np.random.default_rng(3).permutation(P.nelv))
    v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
    f, sens = [_nan(P.n) for _ in range(3)], _nan(P.n)
    op.set_gs_mode(1)
    for _ in range(3):
        op.step(v, ub, f, rho=rho, sens=sens)
    fused, nin, ntot = op.gs_info()
    assert not fused and nin == ntot > 0
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), ref[c]) <= TOL
    assert rel_l2(sens.cpu().numpy(), so) <= TOL
    op.set_gs_mode(0)          # CSR class lists
    g = [_nan(P.n) for _ in range(3)]
    op.step(v, ub, g, rho=rho)
    for c in range(3):
        assert torch.equal(f[c], g[c]), "packed-list and CSR passes must be bit-identical"
    _check_step_host(op, v, ub, rho, f, sens)
    op.free()


def _copies_identical(f, cid):
    """every member of a node class holds the same bits (the C0 property gs_op guarantees)"""
    order = np.argsort(cid, kind="stable")
    fs, cs = f[order], cid[order]
    starts = np.flatnonzero(np.r_[True, cs[1:] != cs[:-1]])
    return np.array_equal(np.maximum.reduceat(fs, starts), np.minimum.reduceat(fs, starts))


@pytest.mark.parametrize("kind", ["box", "ragged", "folded_keys", "x_periodic", "pipe"])
def test_step_xstage(oracle, kind):
    """lx = 8 staged direct-stiffness summation of b200_adjrhs_step.  Level 1: every slot walks a contiguous run of
    elements and sums the x-face pair classes of consecutive elements in the kernel -- BIT-identical to the plain
    kernel + full pass.  Level 2 (default): product classes (faces, 2x2 edges, 2x2x2 vertices) are summed direction
    by direction, x in the kernel, y and z by face passes -- within 1e-12 of the oracle, within rounding of level 0,
    and all copies of a node identical.  Links are only taken where the class lists prove them: meshes with
    unrelated nodes identified, periodic wrap-around, ragged slot runs and the reference's pipe mesh."""
    import os
    from neko_top_b200 import sem, workloads
    lx = 8
    if kind == "pipe":
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "debugging_pipe_mesh.npz")
        m = workloads.load_hex_fixture(path, lx)
        sp = sem.Space(lx)
        x, y, z = m.coords("cpu")
        kt = m.node_keys("cpu")
        G, jac, B = sem.geometric_factors(x, y, z, sp)
        fl = workloads.make_fields(m, x, y, z, kt)
        fnp = lambda a: a.reshape(-1).numpy().copy()

        class P:            # same attributes as helpers.Problem
            pass
        P.lx, P.nelv, P.n, P.D, P.w = lx, m.nelv, m.n, sp.dx, sp.wx
        P.G, P.B, P.v, P.ub, P.rho = [fnp(g) for g in G], fnp(B), [fnp(a) for a in fl.v], [fnp(a) for a in fl.ub], fnp(fl.rho)
        P.cuda = staticmethod(lambda name: ([torch.as_tensor(a).cuda() for a in getattr(P, name)]
                                            if isinstance(getattr(P, name), list) else torch.as_tensor(getattr(P, name)).cuda()))
        keys = fnp(kt).astype(np.int64)
        ne = None
    else:
        ne = {"box": (12, 10, 9), "ragged": (7, 67, 1), "folded_keys": (9, 8, 7), "x_periodic": (16, 8, 5)}[kind]
        P = Problem(lx, ne=ne, deform=0.02)
        keys = P.keys.reshape(-1).numpy().copy()
    if kind == "folded_keys":          # unrelated nodes identified: some face pairs become 3+ member classes
        rng = np.random.default_rng(23)
        sel = rng.random(keys.size) < 0.01
        keys[sel] = keys.max() + 1 + rng.integers(0, sel.sum() // 3 + 1, sel.sum())
    if kind == "x_periodic":           # i = 7 face of the last element of a row glued to i = 0 of the first
        k4 = keys.reshape(ne[2], ne[1], ne[0], lx, lx, lx)       # [ez, ey, ex, k, j, i]
        k4[:, :, -1, :, :, -1] = k4[:, :, 0, :, :, 0]
        keys = k4.reshape(-1)
    fo, so, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho)
    cid, nc = oracle.gs_classes(keys)
    ref = [oracle.gs_add(fo[c], cid, nc) for c in range(3)]
    coef = _ops().coef_t(_ops().space_t(lx, P.D, P.w), P.nelv, P.cuda("G"), P.cuda("B"))
    op = _ops().fused_adjoint_rhs_t(coef)
    op.gs.init(keys)
    v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
    res = {}
    for level in (0, 1, 2):
        op.set_xstage(level)
        f, sens = [_nan(P.n) for _ in range(3)], _nan(P.n)
        for _ in range(2):
            op.step(v, ub, f, rho=rho, sens=sens)
        active, nstaged, nleft, ntot = op.xstage_info()
        assert active == level and nleft == ntot - nstaged
        res[level] = ([a.cpu().numpy() for a in f], sens.cpu().numpy(), nstaged)
        for c in range(3):
            assert rel_l2(res[level][0][c], ref[c]) <= TOL, (kind, level, c)
            assert _copies_identical(res[level][0][c], cid), (kind, level, c)
        assert rel_l2(res[level][1], so) <= TOL
    assert res[0][2] == 0 and 0 <= res[1][2] <= res[2][2] < ntot and res[2][2] > 0
    if kind != "pipe":                # 160 elements on 444 slots: every run is one element, no x links
        assert res[1][2] > 0
    if kind == "box":
        nslots = 444
        # level 1: the 64 pair classes... of every x face inside a run; faces on the domain boundary have more pairs
        assert res[1][2] >= 36 * ((ne[0] - 1) * ne[1] * ne[2] - nslots)
        # level 2: nearly everything is a product class on a box; what is left are the classes cut by run starts
        assert ntot - res[2][2] <= 0.25 * ntot      # 1080 elements on 444 slots: a run start every 2-3 elements
    for c in range(3):
        assert np.array_equal(res[0][0][c], res[1][0][c]), "level 1 must be bit-identical to the plain kernel + full pass"
        scale = np.abs(res[0][0][c]).max()
        assert np.abs(res[2][0][c] - res[0][0][c]).max() <= 4e-15 * scale, "level 2 differs from level 0 by rounding only"
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][1], res[2][1])
    # static forcing (the NF_FULL kernel variant) through the staged path too
    fs = [torch.as_tensor(np.random.default_rng(4).standard_normal(P.n)).cuda() for _ in range(3)]
    out = {}
    for level in (0, 1, 2):
        op.set_xstage(level)
        g3 = [_nan(P.n) for _ in range(3)]
        op.step(v, ub, g3, rho=rho, fstatic=fs)
        out[level] = [a.cpu().numpy() for a in g3]
    for c in range(3):
        assert np.array_equal(out[0][c], out[1][c])
        assert np.abs(out[2][c] - out[0][c]).max() <= 4e-15 * np.abs(out[0][c]).max()
    op.free()


@pytest.mark.parametrize("lx,ne", [(4, (16, 10, 9)), (5, (14, 12, 10)), (6, (12, 10, 8)), (7, (10, 10, 8)),
                                   (9, (10, 8, 6)), (10, (9, 8, 6))])
def test_step_other_orders_irregular_mesh(oracle, lx, ne):
    """b200_adjrhs_step at the other polynomial orders (the pencil kernels) on a mesh with more elements than
    element slots, some unrelated nodes identified and the i = lx-1 face of the last element of every row glued to
    i = 0 of the first: <= 1e-12 against the oracle, all copies of a node identical, static forcing, host-buffer
    step bit-identical.  The staged summation is an lx = 8 feature (measured slower with the pencil kernels,
    profiles/r02C_*): its setting is accepted, reports inactive, and changes no bit."""
    P = Problem(lx, ne=ne, deform=0.02)
    keys = P.keys.reshape(-1).numpy().copy()
    k4 = keys.reshape(ne[2], ne[1], ne[0], lx, lx, lx)
    k4[:, :, -1, :, :, -1] = k4[:, :, 0, :, :, 0]
    keys = k4.reshape(-1).copy()
    rng = np.random.default_rng(29 + lx)
    sel = rng.random(keys.size) < 0.005
    keys[sel] = keys.max() + 1 + rng.integers(0, sel.sum() // 3 + 1, sel.sum())
    fo, so, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho)
    cid, nc = oracle.gs_classes(keys)
    ref = [oracle.gs_add(fo[c], cid, nc) for c in range(3)]
    op, _ = _fused(P)
    op.gs.init(keys)
    gcid, gnc = op.gs.classes()
    assert gnc == nc and np.array_equal(gcid, cid)
    v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
    res = {}
    for level in (0, 2):
        op.set_xstage(level)
        f, sens = [_nan(P.n) for _ in range(3)], _nan(P.n)
        for _ in range(2):
            op.step(v, ub, f, rho=rho, sens=sens)
        active, nstaged, nleft, ntot = op.xstage_info()
        assert active == 0 and nstaged == 0 and nleft == ntot > 0      # (classes with more than one member)
        res[level] = (f, sens)
        for c in range(3):
            assert rel_l2(f[c].cpu().numpy(), ref[c]) <= TOL, (lx, level, c)
            assert _copies_identical(f[c].cpu().numpy(), cid), (lx, level, c)
        assert rel_l2(sens.cpu().numpy(), so) <= TOL
    for c in range(3):
        assert torch.equal(res[0][0][c], res[2][0][c])
    assert torch.equal(res[0][1], res[2][1])
    fs = [torch.as_tensor(np.random.default_rng(4).standard_normal(P.n)).cuda() for _ in range(3)]
    fos, _, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho,
                                   fstatic=[a.cpu().numpy() for a in fs])
    g3 = [_nan(P.n) for _ in range(3)]
    op.step(v, ub, g3, rho=rho, fstatic=fs)
    for c in range(3):
        assert rel_l2(g3[c].cpu().numpy(), oracle.gs_add(fos[c], cid, nc)) <= TOL
    _check_step_host(op, v, ub, rho, res[0][0], res[0][1])
    op.free()


def test_step_gs_packed_lists_irregular_classes(oracle):
    """Node classes of every size (3, 5..16 members and > 16, which stay with the list kernel): keys of a
    box mesh folded modulo a small number so that unrelated nodes are identified."""
    lx = 8
    P = Problem(lx, ne=(9, 8, 7), deform=0.02)
    keys = P.keys.reshape(-1).numpy().copy()
    rng = np.random.default_rng(17)
    sel = rng.random(keys.size) < 0.02
    keys[sel] = keys.max() + 1 + rng.integers(0, 450, sel.sum())       # ~11 members on average, some > 16
    sel2 = (~sel) & (rng.random(keys.size) < 0.01)
    keys[sel2] = keys.max() + 1 + rng.integers(0, sel2.sum() // 3 + 1, sel2.sum())   # ~3 members each
    fo, _, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho)
    cid, nc = oracle.gs_classes(keys)
    counts = np.bincount(cid)
    assert counts.max() > 16 and (counts == 3).any() and (counts == 2).any()
    op, _ = _fused(P)
    op.gs.init(keys)
    op.set_xstage(1)
    gcid, gnc = op.gs.classes()
    assert gnc == nc and np.array_equal(gcid, cid)
    v, ub, rho = P.cuda("v"), P.cuda("ub"), P.cuda("rho")
    f = [_nan(P.n) for _ in range(3)]
    op.set_gs_mode(1)
    op.step(v, ub, f, rho=rho)
    fused, nin, ntot = op.gs_info()
    assert not fused and 0 < nin < ntot
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), oracle.gs_add(fo[c], cid, nc)) <= TOL
    op.set_gs_mode(0)
    g = [_nan(P.n) for _ in range(3)]
    op.step(v, ub, g, rho=rho)
    for c in range(3):
        assert torch.equal(f[c], g[c])
    _check_step_host(op, v, ub, rho, f, None)
    op.free()


@pytest.mark.parametrize("lx,dealias", [(6, False), (8, False), (6, True)])
def test_reference_pipe_mesh(oracle, lx, dealias):
    """SURVEY.md 8d correctness gate: the reference's own mesh fixture data/debugging_pipe.nmsh (160 hexahedra in
    the file's element/vertex numbering; tests/golden/debugging_pipe_mesh.npz): gather-scatter classes exactly
    equal to the oracle's, the fused step (GLL and dealiased operator) within 1e-12."""
    import os
    from neko_top_b200 import sem, workloads
    ops = _ops()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "debugging_pipe_mesh.npz")
    m = workloads.load_hex_fixture(path, lx)
    sp = sem.Space(lx)
    x, y, z = m.coords("cuda")
    keys = m.node_keys("cuda")
    G, _, B = sem.geometric_factors(x, y, z, sp)
    fl = workloads.make_fields(m, x, y, z, keys)
    flat = lambda a: a.reshape(-1).contiguous()
    Gf, Bf, v, ub, rho = [flat(g) for g in G], flat(B), [flat(a) for a in fl.v], [flat(a) for a in fl.ub], flat(fl.rho)
    op = ops.fused_adjoint_rhs_t(ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), m.nelv, Gf, Bf))
    op.gs.init(flat(keys))
    if dealias:
        op.set_dealias(True)
    c = lambda t: t.cpu().numpy()
    cid, nc = oracle.gs_classes(c(flat(keys)))
    gcid, gnc = op.gs.classes()
    assert gnc == nc and np.array_equal(gcid, cid), "gather-scatter index map differs from the oracle's"
    f, sens = [_nan(m.n) for _ in range(3)], _nan(m.n)
    op.step(v, ub, f, rho=rho, sens=sens)
    fo, so, _ = oracle.adjoint_rhs([c(a) for a in v], [c(a) for a in ub], lx, m.nelv, sp.dx, sp.wx,
                                   [c(g) for g in Gf], c(Bf), rho=c(rho), lxd=(3 * lx // 2 if dealias else 0))
    for k in range(3):
        assert rel_l2(c(f[k]), oracle.gs_add(fo[k], cid, nc)) <= TOL
    assert rel_l2(c(sens), so) <= TOL
    op.free()


def test_edge_cases(oracle):
    """Ragged inputs: a single element (no node classes at all), element counts that do not fill the persistent
    grid's slots (1, 2, 5, 443, 445 elements at lx = 8 with 444 slots), and a mesh whose classes are all
    singletons (empty gather-scatter)."""
    ops = _ops()
    lx = 8
    sp_ops = lambda P: ops.fused_adjoint_rhs_t(ops.coef_t(ops.space_t(P.lx, P.D, P.w), P.nelv, P.cuda("G"), P.cuda("B")))
    for ne in ((1, 1, 1), (2, 1, 1), (5, 1, 1), (443, 1, 1), (5, 89, 1)):
        P = Problem(lx, ne=ne, deform=0.0 if ne[0] > 100 else 0.02)
        fo, so, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho)
        cid, nc = oracle.gs_classes(P.keys.reshape(-1).numpy())
        op = sp_ops(P)
        op.gs.init(P.keys.reshape(-1).cuda())
        for mode in (0, 1):
            op.set_gs_mode(mode)
            f, sens = [_nan(P.n) for _ in range(3)], _nan(P.n)
            op.step(P.cuda("v"), P.cuda("ub"), f, rho=P.cuda("rho"), sens=sens)
            for c in range(3):
                assert rel_l2(f[c].cpu().numpy(), oracle.gs_add(fo[c], cid, nc)) <= TOL, (ne, mode)
            assert rel_l2(sens.cpu().numpy(), so) <= TOL
        op.free()
    # all classes singletons: unique keys -> gs is the identity
    P = Problem(lx, ne=(3, 2, 1), deform=0.02)
    op = sp_ops(P)
    op.gs.init(torch.arange(P.n, device="cuda", dtype=torch.int64))
    fo, _, _ = oracle.adjoint_rhs(P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho)
    f = [_nan(P.n) for _ in range(3)]
    op.step(P.cuda("v"), P.cuda("ub"), f, rho=P.cuda("rho"))
    for c in range(3):
        assert rel_l2(f[c].cpu().numpy(), fo[c]) <= TOL
    op.free()
