"""CPU tests of the oracle itself (test infrastructure): known-answer quadrature data, the C oracle
against its independent numpy twin, mathematical identities of the adjoint operator, and the small
golden fixtures under tests/golden/.  The reference tree holds no golden vector for this path
("parity unpinned"), so these identities are what anchors the oracle."""
import json
import os

import numpy as np
import pytest

from helpers import Problem, rel_l2
from oracle import np_oracle as npo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- published known answers: Gauss-Lobatto-Legendre / Gauss-Legendre rules ------------------------
def test_gll_known_answers(oracle):
    z, w = oracle.zwgll(3)
    assert np.allclose(z, [-1, 0, 1], atol=1e-15) and np.allclose(w, [1 / 3, 4 / 3, 1 / 3], atol=1e-15)
    z, w = oracle.zwgll(4)
    assert np.allclose(z, [-1, -1 / np.sqrt(5), 1 / np.sqrt(5), 1], atol=1e-15)
    assert np.allclose(w, [1 / 6, 5 / 6, 5 / 6, 1 / 6], atol=1e-15)
    z, w = oracle.zwgll(5)
    assert np.allclose(z, [-1, -np.sqrt(3 / 7), 0, np.sqrt(3 / 7), 1], atol=1e-15)
    assert np.allclose(w, [1 / 10, 49 / 90, 32 / 45, 49 / 90, 1 / 10], atol=1e-15)


def test_gl_known_answers(oracle):
    z, w = oracle.zwgl(2)
    assert np.allclose(z, [-1 / np.sqrt(3), 1 / np.sqrt(3)], atol=1e-15) and np.allclose(w, [1, 1], atol=1e-15)
    z, w = oracle.zwgl(3)
    assert np.allclose(z, [-np.sqrt(3 / 5), 0, np.sqrt(3 / 5)], atol=1e-15)
    assert np.allclose(w, [5 / 9, 8 / 9, 5 / 9], atol=1e-15)
    for n in (8, 12, 15):
        z, w = oracle.zwgl(n)
        zr, wr = np.polynomial.legendre.leggauss(n)
        assert np.allclose(z, zr, atol=1e-14) and np.allclose(w, wr, atol=1e-14)


@pytest.mark.parametrize("lx", [4, 5, 6, 7, 8, 9, 10])
def test_speclib_properties(oracle, lx):
    z, w = oracle.zwgll(lx)
    zn, wn = npo.gll(lx)
    assert np.allclose(z, zn, atol=1e-14) and np.allclose(w, wn, atol=1e-14)
    assert abs(w.sum() - 2.0) < 1e-14
    # GLL quadrature is exact up to degree 2N-1
    for d in range(0, 2 * lx - 2):
        exact = 0.0 if d % 2 else 2.0 / (d + 1)
        assert abs(np.dot(w, z ** d) - exact) < 1e-13
    D = oracle.dgll(z)
    assert np.allclose(D, npo.dgll(z), atol=1e-12)
    assert np.allclose(D, oracle.deriv_matrix(z), atol=1e-11)   # dgll == generic Lagrange derivative
    for d in range(lx):                                          # exact on polynomials of degree <= N
        assert np.allclose(D @ z ** d, d * z ** max(d - 1, 0) if d else 0 * z, atol=1e-11)
    lxd = 3 * lx // 2
    zd, _ = oracle.zwgl(lxd)
    J = oracle.interp_matrix(zd, z)
    assert np.allclose(J, npo.interp(zd, z), atol=1e-13)
    assert np.allclose(J.sum(axis=1), 1.0, atol=1e-13)
    assert np.allclose(J @ z ** (lx - 1), zd ** (lx - 1), atol=1e-12)


# ---- geometry ---------------------------------------------------------------------------------------
def test_geometry_cofactor_convention(oracle):
    """Cofactors are J-scaled: sum_a G[a,i] * d(x_j)/d(r_a) = jac * delta_ij; B = jac*w3."""
    P = Problem(6, ne=(2, 2, 1), deform=0.04)
    x, y, z = [a.reshape(-1).numpy() for a in P.xyz]
    G, jac, B = oracle.geom(P.lx, P.nelv, P.D, P.w, x, y, z)
    for a, b in zip(G, P.G):
        assert rel_l2(a, b) < 1e-13          # product-side sem.geometric_factors agrees
    assert rel_l2(jac, P.jac) < 1e-13 and rel_l2(B, P.B) < 1e-13
    w3 = np.tile(np.einsum("k,j,i->kji", P.w, P.w, P.w).reshape(-1), P.nelv)
    # weak gradient of the coordinate x itself is B*(1,0,0)
    gx, gy, gz = oracle.opgrad(x, P.lx, P.nelv, P.D, w3[:P.lx ** 3], G)
    assert rel_l2(gx, B) < 1e-12
    assert np.linalg.norm(gy) < 1e-12 * np.linalg.norm(B) and np.linalg.norm(gz) < 1e-12 * np.linalg.norm(B)
    assert jac.min() > 0


# ---- C oracle against the numpy twin ------------------------------------------------------------------
def _v4(P, a):
    return [npo.v4(x, P.lx, P.nelv) for x in a] if isinstance(a, list) else npo.v4(a, P.lx, P.nelv)


@pytest.mark.parametrize("lx", [4, 5, 7, 8])
def test_adjoint_advection_matches_twin(oracle, lx):
    P = Problem(lx, ne=(2, 2, 2), deform=0.03)
    f0 = [np.zeros(P.n) for _ in range(3)]
    fc = oracle.adjoint_advection_no_dealias(f0, P.v, P.ub, lx, P.nelv, P.D, P.w, P.G)
    fn = npo.adjoint_advection(_v4(P, f0), _v4(P, P.v), _v4(P, P.ub), P.D, P.w, _v4(P, P.G))
    for c in range(3):
        assert rel_l2(fc[c], fn[c].reshape(-1)) < 1e-13


@pytest.mark.parametrize("lx,lxd", [(4, 6), (6, 9), (8, 12)])
def test_dealiased_advection_matches_twin(oracle, lx, lxd):
    P = Problem(lx, ne=(2, 1, 2), deform=0.03)
    f0 = [np.zeros(P.n) for _ in range(3)]
    fc = oracle.adjoint_advection_dealias(f0, P.v, P.ub, lx, lxd, P.nelv, P.G)
    fn = npo.adjoint_advection_dealias(_v4(P, f0), _v4(P, P.v), _v4(P, P.ub), lx, lxd, _v4(P, P.G))
    for c in range(3):
        assert rel_l2(fc[c], fn[c].reshape(-1)) < 1e-12


def test_composite_matrix_formulation_of_the_dealiased_operator(oracle):
    """The formulation the tensor-core kernel evaluates (csrc/advop_mma_kernel.cuh), written out in numpy: with
    DJ = D_fine J, forward T = (J x J x J) q and d_r U = (DJ x J x J) U ..., backward out = (J^T x J^T x J^T) R +
    (DJ^T x J^T x J^T) Fr + (J^T x DJ^T x J^T) Fs + (J^T x J^T x DJ^T) Ft.  Must equal the oracle's dealiased adjoint
    operator (interpolate, opgrad on the fine grid, cdtp, project: adv_adjoint_dealias.f90:358-456) to rounding."""
    from neko_top_b200 import sem
    lx, lxd = 8, 12
    P = Problem(lx, ne=(2, 1, 2), deform=0.03)
    ds = sem.DealiasSpace(lx, lxd)
    J, DJ, wd = ds.interp, ds.dxd @ ds.interp, ds.wd
    A = lambda a: np.asarray(a, dtype=np.float64).reshape(P.nelv, lx, lx, lx)              # [e, k, j, i]
    ap = lambda Mr, Ms, Mt, q: np.einsum("ai,bj,ck,ekji->ecba", Mr, Ms, Mt, q, optimize=True)   # fine [e, c, b, a]
    bk = lambda Mr, Ms, Mt, q: np.einsum("ai,bj,ck,ecba->ekji", Mr, Ms, Mt, q, optimize=True)   # transposed
    v, ub, G = [A(a) for a in P.v], [A(a) for a in P.ub], [ap(J, J, J, A(g)) for g in P.G]
    tv, tb = [ap(J, J, J, a) for a in v], [ap(J, J, J, a) for a in ub]
    dU = [(ap(DJ, J, J, a), ap(J, DJ, J, a), ap(J, J, DJ, a)) for a in ub]
    w3 = np.einsum("c,b,a->cba", wd, wd, wd)[None]
    R = [sum(tv[c] * w3 * (G[3 * d] * dU[c][0] + G[3 * d + 1] * dU[c][1] + G[3 * d + 2] * dU[c][2]) for c in range(3))
         for d in range(3)]
    cr, cs, ct = (w3 * sum(tb[k] * G[3 * k + e] for k in range(3)) for e in range(3))
    rng = np.random.default_rng(12)
    f0 = [rng.standard_normal(P.n) for _ in range(3)]
    fo = oracle.adjoint_advection_dealias(f0, P.v, P.ub, lx, lxd, P.nelv, P.G)
    for c in range(3):
        out = bk(J, J, J, R[c]) + bk(DJ, J, J, tv[c] * cr) + bk(J, DJ, J, tv[c] * cs) + bk(J, J, DJ, tv[c] * ct)
        assert rel_l2(f0[c] - out.reshape(-1), fo[c]) < 1e-13


def test_full_rhs_matches_twin(oracle):
    P = Problem(6, ne=(2, 2, 2), deform=0.03)
    rng = np.random.default_rng(7)
    fs = [rng.standard_normal(P.n) for _ in range(3)]
    fc, sc, chi = oracle.adjoint_rhs(P.v, P.ub, P.lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho, fstatic=fs)
    chin = npo.ramp(P.rho)
    assert np.array_equal(chi, chin)
    fn = npo.adjoint_rhs(_v4(P, P.v), _v4(P, P.ub), P.D, P.w, _v4(P, P.G), _v4(P, P.B), _v4(P, chin),
                         fstatic=_v4(P, fs))
    for c in range(3):
        assert rel_l2(fc[c], fn[c].reshape(-1)) < 1e-13
    assert rel_l2(sc, npo.sensitivity(P.ub, P.v).reshape(-1)) < 1e-14


# ---- mathematical identities that anchor the operator ---------------------------------------------------
@pytest.mark.parametrize("lx", [5, 8])
def test_discrete_adjoint_identity(oracle, lx):
    """sum_i <w_i, L(v)_i> == sum_i <A(w)_i, v_i> to round-off, L = compute_linear
    (adv_adjoint_no_dealias.f90:404-424), A = compute_adjoint (:162-201), on a non-affine mesh."""
    P = Problem(lx, ne=(2, 2, 2), deform=0.05)
    rng = np.random.default_rng(3)
    wv = [rng.standard_normal(P.n) for _ in range(3)]
    z = [np.zeros(P.n) for _ in range(3)]
    Lv = oracle.linear_advection_no_dealias(z, P.v, P.ub, lx, P.nelv, P.D, P.w, P.G, P.jac)
    Aw = oracle.adjoint_advection_no_dealias(z, wv, P.ub, lx, P.nelv, P.D, P.w, P.G)
    lhs = sum(np.dot(wv[c], Lv[c]) for c in range(3))
    rhs = sum(np.dot(Aw[c], P.v[c]) for c in range(3))
    scale = np.sqrt(sum(np.dot(a, a) for a in wv)) * np.sqrt(sum(np.dot(a, a) for a in Lv))
    assert abs(lhs - rhs) / scale < 1e-12


def test_discrete_adjoint_identity_dealias(oracle):
    lx, lxd = 6, 9
    P = Problem(lx, ne=(2, 2, 1), deform=0.05)
    rng = np.random.default_rng(4)
    wv = [rng.standard_normal(P.n) for _ in range(3)]
    z = [np.zeros(P.n) for _ in range(3)]
    Lv = oracle.linear_advection_dealias(z, P.v, P.ub, lx, lxd, P.nelv, P.G)
    Aw = oracle.adjoint_advection_dealias(z, wv, P.ub, lx, lxd, P.nelv, P.G)
    lhs = sum(np.dot(wv[c], Lv[c]) for c in range(3))
    rhs = sum(np.dot(Aw[c], P.v[c]) for c in range(3))
    scale = np.sqrt(sum(np.dot(a, a) for a in wv)) * np.sqrt(sum(np.dot(a, a) for a in Lv))
    assert abs(lhs - rhs) / scale < 1e-12


def test_polynomial_exactness(oracle):
    """Affine element, U_b linear in x: (grad U_b)^T v term is B * (dU_j/dx_i) v_j exactly."""
    lx = 6
    P = Problem(lx, ne=(2, 1, 1), deform=0.0)
    x, y, z = [a.reshape(-1).numpy() for a in P.xyz]
    ub = [2.0 * x + 3.0 * y - z, 0.5 * y + x, -1.5 * z + 0.25 * y]
    grad = np.array([[2.0, 3.0, -1.0], [1.0, 0.5, 0.0], [0.0, 0.25, -1.5]])   # grad[j][i] = dU_j/dx_i
    w3 = np.einsum("k,j,i->kji", P.w, P.w, P.w).reshape(-1)
    for j in range(3):
        g = oracle.opgrad(ub[j], lx, P.nelv, P.D, w3, P.G)
        for i in range(3):
            assert np.allclose(g[i], grad[j][i] * P.B, atol=1e-13)
    # cdtp is the exact transpose of the weak derivative: <cdtp(x; dr..), y> == <x, w3*(dr y_r + ds y_s + dt y_t)>
    rng = np.random.default_rng(5)
    a, b = rng.standard_normal(P.n), rng.standard_normal(P.n)
    lhs = np.dot(oracle.cdtp(a, P.G[0], P.G[1], P.G[2], lx, P.nelv, P.D, w3), b)
    rhs = np.dot(a, oracle.opgrad(b, lx, P.nelv, P.D, w3, P.G)[0])
    assert abs(lhs - rhs) < 1e-12 * np.linalg.norm(a) * np.linalg.norm(b)


def test_linearity_in_adjoint_velocity(oracle):
    P = Problem(5, ne=(2, 2, 1), deform=0.03)
    Q = Problem(5, ne=(2, 2, 1), deform=0.03, seed_shift=977)
    z = [np.zeros(P.n) for _ in range(3)]
    a, b = 0.7, -1.3
    A1 = oracle.adjoint_advection_no_dealias(z, P.v, P.ub, 5, P.nelv, P.D, P.w, P.G)
    A2 = oracle.adjoint_advection_no_dealias(z, Q.v, P.ub, 5, P.nelv, P.D, P.w, P.G)
    mix = [a * P.v[c] + b * Q.v[c] for c in range(3)]
    Am = oracle.adjoint_advection_no_dealias(z, mix, P.ub, 5, P.nelv, P.D, P.w, P.G)
    for c in range(3):
        assert rel_l2(Am[c], a * A1[c] + b * A2[c]) < 1e-13


def test_bug_compat_is_observable(oracle):
    """Defects D1 (index shift) and D2 (sign) of the reference CPU branch change the result by O(1);
    the oracle's default is the intended (device-branch) operator."""
    P = Problem(5, ne=(2, 1, 1), deform=0.02)
    z = [np.zeros(P.n) for _ in range(3)]
    good = oracle.adjoint_advection_no_dealias(z, P.v, P.ub, 5, P.nelv, P.D, P.w, P.G, bug_compat=0)
    bad = oracle.adjoint_advection_no_dealias(z, P.v, P.ub, 5, P.nelv, P.D, P.w, P.G, bug_compat=1)
    assert rel_l2(bad[0], good[0]) > 0.1


# ---- pointwise terms ------------------------------------------------------------------------------------
def test_ramp_forward_backward(oracle):
    rho = np.linspace(0.0, 1.0, 101)
    for cu in (1, 0):
        chi = oracle.ramp(rho, 0.0, 1000.0, 1.0, cu)
        assert chi[0] == 0.0 and abs(chi[-1] - 1000.0) < 1e-12
        assert np.all(np.diff(chi) > 0)
        h = 1e-6
        fd = (oracle.ramp(rho + h, 0.0, 1000.0, 1.0, cu) - oracle.ramp(rho - h, 0.0, 1000.0, 1.0, cu)) / (2 * h)
        an = oracle.ramp_backward(np.ones_like(rho), rho, 0.0, 1000.0, 1.0, cu)
        assert np.allclose(an, fd, rtol=1e-6)
    # convex-up lies above the chord, convex-down below
    mid = oracle.ramp(np.array([0.5]), 0.0, 1000.0, 1.0, 1)[0]
    assert mid > 500.0 > oracle.ramp(np.array([0.5]), 0.0, 1000.0, 1.0, 0)[0]


def test_source_terms_and_mask(oracle):
    rng = np.random.default_rng(11)
    n = 500
    u = [rng.standard_normal(n) for _ in range(3)]
    chi = rng.random(n) * 1000
    f = [rng.standard_normal(n) for _ in range(3)]
    fb = oracle.brinkman(f, u, chi)
    for c in range(3):
        assert np.array_equal(fb[c], f[c] - u[c] * chi)
    mask = np.sort(rng.choice(n, 77, replace=False)).astype(np.int32) + 1      # 1-based
    fl = oracle.lube(f, u, chi, 2.5, mask)
    keep = np.zeros(n)
    keep[mask - 1] = chi[mask - 1] * 2.5
    for c in range(3):
        assert np.array_equal(fl[c], f[c] + u[c] * keep)
    fo = oracle.opcolv(f, chi)
    assert np.array_equal(fo[1], f[1] * chi)
    S = oracle.sensitivity(u, f, 1.0, 1)
    assert np.allclose(S, -(u[0] * f[0] + u[1] * f[1] + u[2] * f[2]) + (u[0] ** 2 + u[1] ** 2 + u[2] ** 2), atol=1e-12)


# ---- gather-scatter ------------------------------------------------------------------------------------
def test_gs_classes_and_add(oracle):
    P = Problem(4, ne=(3, 2, 2), deform=0.0)
    key = P.keys.reshape(-1).numpy()
    cid, nc = oracle.gs_classes(key)
    lx = 4
    assert nc == (3 * (lx - 1) + 1) * (2 * (lx - 1) + 1) * (2 * (lx - 1) + 1)
    # same partition as the keys, canonical labels (first appearance)
    assert np.array_equal(cid[np.unique(cid, return_index=True)[1]], np.arange(nc))
    for a, b in ((0, 1), (5, 100), (17, 333)):
        assert (key[a] == key[b]) == (cid[a] == cid[b])
    _, inv = np.unique(key, return_inverse=True)
    pairs = set(zip(cid.tolist(), inv.tolist()))
    assert len(pairs) == nc
    rng = np.random.default_rng(2)
    f = rng.standard_normal(P.n)
    g = oracle.gs_add(f, cid, nc)
    assert rel_l2(g, npo.gs_add(f, key)) < 1e-14
    # multiplicity: gs(1) counts members; gs is idempotent after weighting by 1/mult
    mult = oracle.gs_add(np.ones(P.n), cid, nc)
    assert mult.min() == 1 and mult.max() == 8
    g2 = oracle.gs_add(g / mult, cid, nc)
    assert rel_l2(g2, g) < 1e-14
    # sum of all entries weighted by 1/mult is conserved
    assert abs((g / mult).sum() - f.sum()) < 1e-10


# ---- golden fixtures --------------------------------------------------------------------------------------
def test_golden_fixture(oracle):
    """tests/golden/adjrhs_lx4.json was produced by tests/golden/make_golden.py from the numpy twin; it
    pins both oracles against silent regressions (it is NOT reference output: parity unpinned)."""
    with open(os.path.join(GOLDEN, "adjrhs_lx4.json")) as fh:
        g = json.load(fh)
    P = Problem(g["lx"], ne=tuple(g["ne"]), deform=g["deform"])
    f, s, chi = oracle.adjoint_rhs(P.v, P.ub, P.lx, P.nelv, P.D, P.w, P.G, P.B, rho=P.rho)
    idx = np.asarray(g["idx"])
    for c in range(3):
        assert np.allclose(f[c][idx], g["f"][c], rtol=1e-12, atol=1e-13)
    assert np.allclose(s[idx], g["sens"], rtol=1e-13, atol=1e-14)
    assert np.allclose(chi[idx], g["chi"], rtol=1e-14)
    assert abs(sum(np.abs(a).sum() for a in f) - g["sum_abs_f"]) < 1e-10 * g["sum_abs_f"]
    fd = oracle.adjoint_advection_dealias([np.zeros(P.n)] * 3, P.v, P.ub, P.lx, g["lxd"], P.nelv, P.G)
    for c in range(3):
        assert np.allclose(fd[c][idx], g["f_dealias"][c], rtol=1e-11, atol=1e-13)


def test_time_scheme_restatement(oracle):
    """sumab / makeabf / makebdf (Neko rhs_maker, restated in oracle.c) against their defining formulas in
    numpy, incl. the in-place lag rotation of makeabf and the order-dependent number of BDF terms."""
    rng = np.random.default_rng(4)
    n = 1001
    r3 = lambda: [rng.standard_normal(n) for _ in range(3)]
    u, l1, l2, f, a1, a2 = r3(), r3(), r3(), r3(), r3(), r3()
    B = rng.random(n) + 0.5
    rho, dt = 0.9, 0.02
    for order, ab, bd in ((2, [2.0, -1.0, 0.0], [1.5, 2.0, -0.5, 0.0]),
                          (3, [3.0, -3.0, 1.0], [11.0 / 6.0, 3.0, -1.5, 1.0 / 3.0])):
        ue = oracle.sumab(u, l1, l2, ab, order)
        for c in range(3):
            ref = ab[0] * u[c] + ab[1] * l1[c] + (ab[2] * l2[c] if order == 3 else 0.0)
            assert np.allclose(ue[c], ref, rtol=1e-15, atol=1e-15)
        n1, n2, nf = oracle.makeabf(a1, a2, f, rho, ab)
        for c in range(3):
            assert np.array_equal(n2[c], a1[c]) and np.array_equal(n1[c], f[c])
            assert np.allclose(nf[c], (ab[0] * f[c] + ab[1] * a1[c] + ab[2] * a2[c]) * rho, rtol=1e-15, atol=1e-15)
        g = oracle.makebdf(l1, l2, f, u, B, rho, dt, bd, order)
        for c in range(3):
            tb = u[c] * B * bd[1] + l1[c] * B * bd[2] + (l2[c] * B * bd[3] if order == 3 else 0.0)
            assert np.allclose(g[c], f[c] + tb * (rho / dt), rtol=1e-14, atol=1e-14)


def test_objective_chain_restatement(oracle):
    """Neko curl / dudxyz / objective restated in oracle.c against closed forms on an affine mesh: the curl of a
    rigid rotation is (0,0,2), curl(curl(u)) of a divergence-free harmonic-like field is -laplace(u), the
    dissipation of u = (sin 2y, 0, 0) on the unit cube is 4(1/2 + sin(4)/8); mask semantics of
    mask_exterior_const / glsc2_mask (1-based indices)."""
    from helpers import Problem
    lx = 8
    P = Problem(lx, ne=(2, 2, 2), deform=0.0)
    x, y, z = [a.reshape(-1).numpy() for a in P.xyz]
    cid, nc = oracle.gs_classes(P.keys.reshape(-1).numpy())
    Binv, jacinv = 1.0 / oracle.gs_add(P.B, cid, nc), 1.0 / P.jac
    w = oracle.curl([-y, x, 0 * x], lx, P.nelv, P.D, P.G, jacinv, P.B, Binv, cid, nc)
    assert np.abs(w[0]).max() < 1e-12 and np.abs(w[1]).max() < 1e-12 and np.abs(w[2] - 2.0).max() < 1e-12
    # u = (sin(y) , 0, 0): div u = 0, curl curl u = -laplace u = (sin y, 0, 0)
    f = oracle.curlcurl_forcing([0 * x] * 3, [np.sin(y), 0 * x, 0 * x], lx, P.nelv, P.D, P.G, jacinv, P.B, Binv,
                                cid, nc, obj_scale=2.0)
    assert np.abs(f[0] - 2.0 * np.sin(y)).max() < 1e-6 and np.abs(f[1]).max() < 1e-8 and np.abs(f[2]).max() < 1e-8
    val, dis, lube = oracle.min_dissipation_objective([np.sin(2 * y), 0 * x, 0 * x], None, lx, P.nelv, P.D, P.G,
                                                      jacinv, P.B, obj_scale=3.0)
    exact = 4.0 * (0.5 + np.sin(4.0) / 8.0)
    assert abs(dis - exact) < 1e-9 and lube == 0.0 and abs(val - 3.0 * dis) < 1e-14
    mask = np.array([1, 5, P.n], dtype=np.int32)
    g = oracle.mask_exterior_const(x + 1.0, mask, -1.0)
    assert g[0] == x[0] + 1.0 and g[4] == x[4] + 1.0 and g[-1] == x[-1] + 1.0 and np.all(np.delete(g, [0, 4, P.n - 1]) == -1.0)


def test_helmholtz_operator_restatement(oracle):
    """Neko ax_helm restated in oracle.c: the assembled operator r^2 K + M on a deformed mesh is symmetric
    positive definite, K annihilates constants (A 1 = M 1), <u, K u> equals the Dirichlet integral of a known
    field on an affine mesh, and the dense PDE-filter solve leaves constants fixed and contracts the range."""
    from helpers import Problem
    lx = 5
    P = Problem(lx, ne=(2, 2, 1), deform=0.04)
    cid, nc = oracle.gs_classes(P.keys.reshape(-1).numpy())
    jacinv = 1.0 / P.jac
    rep = np.unique(cid, return_index=True)[1]
    A = np.zeros((nc, nc))
    for k in range(nc):
        e = (cid == k).astype(np.float64)
        A[:, k] = oracle.gs_add(oracle.ax_helm(e, lx, P.nelv, P.D, P.w, P.G, jacinv, P.B, 0.3, 1.0), cid, nc)[rep]
    assert np.abs(A - A.T).max() <= 1e-12 * np.abs(A).max()
    assert np.linalg.eigvalsh(0.5 * (A + A.T)).min() > 0.0
    one = np.ones(P.n)
    a1 = oracle.ax_helm(one, lx, P.nelv, P.D, P.w, P.G, jacinv, P.B, 0.3, 1.0)
    assert np.abs(a1 - P.B).max() <= 1e-12
    Q = Problem(7, ne=(2, 2, 2), deform=0.0)
    x, y, z = [a.reshape(-1).numpy() for a in Q.xyz]
    u = np.sin(2.0 * x) + y * z
    Ku = oracle.ax_helm(u, 7, Q.nelv, Q.D, Q.w, Q.G, 1.0 / Q.jac, Q.B, 1.0, 0.0)
    exact = (2.0 + 0.5 * np.sin(4.0)) + 1.0 / 3.0 + 1.0 / 3.0      # int 4cos^2(2x) + z^2 + y^2 over the unit cube
    assert abs(float(u @ Ku) - exact) <= 1e-7
    xf = oracle.pde_filter_dense(P.rho, lx, P.nelv, P.D, P.w, P.G, jacinv, P.B, cid, nc, 0.1)
    assert P.rho.min() < xf.min() and xf.max() < P.rho.max()
    assert np.abs(oracle.pde_filter_dense(one, lx, P.nelv, P.D, P.w, P.G, jacinv, P.B, cid, nc, 0.1) - 1.0).max() <= 1e-12


def test_steady_simcomp_restatement():
    """steady_simcomp.f90:154-188: squared norm of the change per field, copy-or-freeze."""
    from oracle import np_oracle as npo
    rng = np.random.default_rng(11)
    new = [rng.standard_normal(1000) for _ in range(4)]
    old = [a + 1e-3 * rng.standard_normal(1000) for a in new]
    expect = [float(np.sum((o - a) ** 2)) for a, o in zip(new, old)]
    nd, freeze = npo.steady_simcomp_compute(new, old, tol=1e-9)
    assert not freeze and np.allclose(nd, expect, rtol=1e-13)
    assert all(np.array_equal(o, a) for a, o in zip(new, old)), "old fields must be overwritten by the new ones"
    nd, freeze = npo.steady_simcomp_compute(new, old, tol=1e-9)
    assert freeze and max(nd) == 0.0
