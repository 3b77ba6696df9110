"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)


def build(force=False):
    """Compile the C oracle in place (gcc; seconds)."""
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.check_call(["make", "-B", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


class params(C.Structure):
    _fields_ = [("f_min", C.c_double), ("f_max", C.c_double), ("q", C.c_double),
                ("convex_up", C.c_int), ("if_lube", C.c_int), ("K_lube", C.c_double),
                ("K_sens", C.c_double), ("lxd", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_gs_classes.restype = C.c_int64
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _G(G):
    arr = (_dp * 9)(*[_p(g) for g in G])
    return arr


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    """OpenMP threads of the element loops (bench.py: all host cores, whatever OMP_NUM_THREADS says)."""
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


# ---- speclib ----
def zwgll(n):
    z, w = np.zeros(n), np.zeros(n)
    lib().orc_zwgll(_p(z), _p(w), C.c_int(n))
    return z, w


def zwgl(n):
    z, w = np.zeros(n), np.zeros(n)
    lib().orc_zwgl(_p(z), _p(w), C.c_int(n))
    return z, w


def dgll(z):
    """Returns D as a numpy (n,n) array with D[i,j] = D(i,j)."""
    n = len(z)
    D = np.zeros(n * n)
    lib().orc_dgll(_p(D), _p(f64(z)), C.c_int(n))
    return D.reshape(n, n).T.copy()      # col-major storage -> D[i,j]


def deriv_matrix(z):
    n = len(z)
    D = np.zeros(n * n)
    lib().orc_deriv_matrix(_p(D), _p(f64(z)), C.c_int(n))
    return D.reshape(n, n).T.copy()


def interp_matrix(zto, zfrom):
    nt, nf = len(zto), len(zfrom)
    J = np.zeros(nt * nf)
    lib().orc_interp_matrix(_p(J), _p(f64(zto)), C.c_int(nt), _p(f64(zfrom)), C.c_int(nf))
    return J.reshape(nf, nt).T.copy()    # J[a,m]


def _colmajor(M):
    """numpy M[i,j] -> flat col-major buffer."""
    return f64(np.asarray(M).T.reshape(-1))


# ---- geometry ----
def geom(lx, nelv, D, w, x, y, z):
    n = lx ** 3 * nelv
    G = [np.zeros(n) for _ in range(9)]
    jac, B = np.zeros(n), np.zeros(n)
    lib().orc_geom(C.c_int(lx), C.c_int(nelv), _p(_colmajor(D)), _p(f64(w)), _p(f64(x)), _p(f64(y)),
                   _p(f64(z)), _G(G), _p(jac), _p(B))
    return G, jac, B


# ---- operators ----
def opgrad(u, lx, nelv, D, w3, G):
    n = lx ** 3 * nelv
    ux, uy, uz = np.zeros(n), np.zeros(n), np.zeros(n)
    lib().orc_opgrad(_p(ux), _p(uy), _p(uz), _p(f64(u)), C.c_int(lx), C.c_int(nelv),
                     _p(_colmajor(D)), _p(f64(w3)), _G(G))
    return ux, uy, uz


def cdtp(x, dr, ds, dt, lx, nelv, D, w3):
    out = np.zeros(lx ** 3 * nelv)
    lib().orc_cdtp(_p(out), _p(f64(x)), _p(f64(dr)), _p(f64(ds)), _p(f64(dt)), C.c_int(lx),
                   C.c_int(nelv), _p(_colmajor(D)), _p(f64(w3)))
    return out


def tnsr3d(u, nu, A, nelv):
    """v = (A x A x A) u with A[a,l] of shape (nv,nu)."""
    nv = A.shape[0]
    v = np.zeros(nv ** 3 * nelv)
    lib().orc_tnsr3d(_p(v), C.c_int(nv), _p(f64(u)), C.c_int(nu), _p(_colmajor(A)), C.c_int(nelv))
    return v


def adjoint_advection_no_dealias(f, v, vb, lx, nelv, D, w, G, bug_compat=0):
    """f, v, vb: lists of 3 arrays; f is accumulated in place (copies returned)."""
    f = [f64(a).copy() for a in f]
    lib().orc_adjoint_advection_no_dealias(_p(f[0]), _p(f[1]), _p(f[2]), _p(f64(v[0])), _p(f64(v[1])),
                                           _p(f64(v[2])), _p(f64(vb[0])), _p(f64(vb[1])), _p(f64(vb[2])),
                                           C.c_int(lx), C.c_int(nelv), _p(_colmajor(D)), _p(f64(w)),
                                           _G(G), C.c_int(bug_compat))
    return f


def linear_advection_no_dealias(f, v, vb, lx, nelv, D, w, G, jac):
    f = [f64(a).copy() for a in f]
    lib().orc_linear_advection_no_dealias(_p(f[0]), _p(f[1]), _p(f[2]), _p(f64(v[0])), _p(f64(v[1])),
                                          _p(f64(v[2])), _p(f64(vb[0])), _p(f64(vb[1])), _p(f64(vb[2])),
                                          C.c_int(lx), C.c_int(nelv), _p(_colmajor(D)), _p(f64(w)),
                                          _G(G), _p(f64(jac)))
    return f


def adjoint_advection_dealias(f, v, vb, lx, lxd, nelv, G):
    f = [f64(a).copy() for a in f]
    lib().orc_adjoint_advection_dealias(_p(f[0]), _p(f[1]), _p(f[2]), _p(f64(v[0])), _p(f64(v[1])),
                                        _p(f64(v[2])), _p(f64(vb[0])), _p(f64(vb[1])), _p(f64(vb[2])),
                                        C.c_int(lx), C.c_int(lxd), C.c_int(nelv), _G(G))
    return f


def linear_advection_dealias(f, v, vb, lx, lxd, nelv, G):
    f = [f64(a).copy() for a in f]
    lib().orc_linear_advection_dealias(_p(f[0]), _p(f[1]), _p(f[2]), _p(f64(v[0])), _p(f64(v[1])),
                                       _p(f64(v[2])), _p(f64(vb[0])), _p(f64(vb[1])), _p(f64(vb[2])),
                                       C.c_int(lx), C.c_int(lxd), C.c_int(nelv), _G(G))
    return f


def ramp(rho, f_min=0.0, f_max=1000.0, q=1.0, convex_up=1):
    chi = np.zeros_like(rho)
    lib().orc_ramp(_p(chi), _p(f64(rho)), C.c_int64(rho.size), C.c_double(f_min), C.c_double(f_max),
                   C.c_double(q), C.c_int(convex_up))
    return chi


def ramp_backward(dF_dchi, rho, f_min=0.0, f_max=1000.0, q=1.0, convex_up=1):
    out = np.zeros_like(rho)
    lib().orc_ramp_backward(_p(out), _p(f64(dF_dchi)), _p(f64(rho)), C.c_int64(rho.size),
                            C.c_double(f_min), C.c_double(f_max), C.c_double(q), C.c_int(convex_up))
    return out


def brinkman(f, u, chi):
    f = [f64(a).copy() for a in f]
    lib().orc_brinkman(_p(f[0]), _p(f[1]), _p(f[2]), _p(f64(u[0])), _p(f64(u[1])), _p(f64(u[2])),
                       _p(f64(chi)), C.c_int64(chi.size))
    return f


def lube(f, u, chi, K, mask=None):
    f = [f64(a).copy() for a in f]
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.int32)
    lib().orc_lube(_p(f[0]), _p(f[1]), _p(f[2]), _p(f64(u[0])), _p(f64(u[1])), _p(f64(u[2])),
                   _p(f64(chi)), C.c_double(K),
                   mask.ctypes.data_as(_ip) if mask is not None else None,
                   C.c_int(0 if mask is None else mask.size), C.c_int64(chi.size))
    return f


def opcolv(f, B):
    f = [f64(a).copy() for a in f]
    lib().orc_opcolv(_p(f[0]), _p(f[1]), _p(f[2]), _p(f64(B)), C.c_int64(B.size))
    return f


def sensitivity(u, ua, K_obj=1.0, if_lube=1):
    S = np.zeros_like(u[0])
    lib().orc_sensitivity(_p(S), _p(f64(u[0])), _p(f64(u[1])), _p(f64(u[2])), _p(f64(ua[0])),
                          _p(f64(ua[1])), _p(f64(ua[2])), C.c_double(K_obj), C.c_int(if_lube),
                          C.c_int64(S.size))
    return S


def adjoint_rhs(v, vb, lx, nelv, D, w, G, B, rho=None, chi=None, fstatic=None, mask=None,
                f_min=0.0, f_max=1000.0, q=1.0, convex_up=1, if_lube=1, K_lube=1.0, K_sens=1.0,
                lxd=0, want_sens=True):
    """Full RHS slice (adjoint_pnpn.f90:661-682).  Returns (f[3], sens, chi)."""
    n = lx ** 3 * nelv
    f = [np.zeros(n) for _ in range(3)]
    sens = np.zeros(n) if want_sens else None
    chi_out = np.zeros(n)
    p = params(f_min, f_max, q, convex_up, if_lube, K_lube, K_sens, lxd)
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.int32)
    fs = [None] * 3 if fstatic is None else [f64(a) for a in fstatic]
    lib().orc_adjoint_rhs(_p(f[0]), _p(f[1]), _p(f[2]), _p(sens), _p(chi_out),
                          _p(f64(v[0])), _p(f64(v[1])), _p(f64(v[2])),
                          _p(f64(vb[0])), _p(f64(vb[1])), _p(f64(vb[2])),
                          _p(f64(rho)) if rho is not None else None,
                          _p(f64(chi)) if chi is not None else None,
                          _p(fs[0]), _p(fs[1]), _p(fs[2]),
                          mask.ctypes.data_as(_ip) if mask is not None else None,
                          C.c_int(0 if mask is None else mask.size),
                          C.c_int(lx), C.c_int(nelv), _p(_colmajor(D)), _p(f64(w)), _G(G), _p(f64(B)),
                          C.byref(p))
    return f, sens, chi_out


# ---- minimum-dissipation objective chain ----
def _mask(mask):
    if mask is None:
        return None, 0
    m = np.ascontiguousarray(mask, dtype=np.int32)
    return m, m.size


def curl(u, lx, nelv, D, G, jacinv, B, Binv, cid, nclass):
    n = lx ** 3 * nelv
    w = [np.zeros(n) for _ in range(3)]
    lib().orc_curl(*[_p(a) for a in w], *[_p(f64(a)) for a in u], C.c_int(lx), C.c_int(nelv), _p(_colmajor(D)),
                   _G(G), _p(f64(jacinv)), _p(f64(B)), _p(f64(Binv)), cid.ctypes.data_as(_lp), C.c_int64(nclass))
    return w


def curlcurl_forcing(f, u, lx, nelv, D, G, jacinv, B, Binv, cid, nclass, mask=None, obj_scale=1.0):
    f = [f64(a).copy() for a in f]
    m, ms = _mask(mask)
    lib().orc_curlcurl_forcing(*[_p(a) for a in f], *[_p(f64(a)) for a in u], C.c_int(lx), C.c_int(nelv),
                               _p(_colmajor(D)), _G(G), _p(f64(jacinv)), _p(f64(B)), _p(f64(Binv)),
                               cid.ctypes.data_as(_lp), C.c_int64(nclass),
                               m.ctypes.data_as(_ip) if m is not None else None, C.c_int(ms), C.c_double(obj_scale))
    return f


def min_dissipation_objective(u, chi, lx, nelv, D, G, jacinv, B, mask=None, K=1.0, obj_scale=1.0):
    """returns (objective, dissipation, lube_value)."""
    out = (C.c_double * 2)()
    m, ms = _mask(mask)
    lib().orc_min_dissipation_objective.restype = C.c_double
    val = lib().orc_min_dissipation_objective(out, *[_p(f64(a)) for a in u], _p(f64(chi)) if chi is not None else None,
                                              C.c_int(lx), C.c_int(nelv), _p(_colmajor(D)), _G(G), _p(f64(jacinv)),
                                              _p(f64(B)), m.ctypes.data_as(_ip) if m is not None else None,
                                              C.c_int(ms), C.c_double(K), C.c_double(obj_scale))
    return val, out[0], out[1]


def mask_exterior_const(fld, mask, c):
    fld = f64(fld).copy()
    m, ms = _mask(mask)
    lib().orc_mask_exterior_const(_p(fld), m.ctypes.data_as(_ip), C.c_int(ms), C.c_double(c), C.c_int64(fld.size))
    return fld


def ax_helm(u, lx, nelv, D, w, G, jacinv, B, h1, h2):
    out = np.zeros(lx ** 3 * nelv)
    lib().orc_ax_helm(_p(out), _p(f64(u)), C.c_int(lx), C.c_int(nelv), _p(_colmajor(D)), _p(f64(w)), _G(G),
                      _p(f64(jacinv)), _p(f64(B)), C.c_double(h1), C.c_double(h2))
    return out


def pde_filter_dense(x_in, lx, nelv, D, w, G, jacinv, B, cid, nclass, radius):
    """Reference solution of the PDE filter on a SMALL mesh: assemble (r^2 K + M) on the unique nodes by applying
    ax_helm + gs to unit vectors, solve with LAPACK.  Returns x on the local dofs."""
    n = lx ** 3 * nelv
    A = np.zeros((nclass, nclass))
    for k in range(nclass):
        e = (cid == k).astype(np.float64)
        col = gs_add(ax_helm(e, lx, nelv, D, w, G, jacinv, B, radius * radius, 1.0), cid, nclass)
        # one representative dof per class
        A[:, k] = col[np.unique(cid, return_index=True)[1]]
    rhs = gs_add(f64(x_in) * f64(B), cid, nclass)[np.unique(cid, return_index=True)[1]]
    xg = np.linalg.solve(A, rhs)
    return xg[cid]


# ---- explicit time scheme (Neko rhs_maker) ----
def sumab(u, ulag1, ulag2, ab, nab):
    out = [np.zeros_like(f64(u[0])) for _ in range(3)]
    abv = (C.c_double * 3)(*[float(a) for a in ab])
    lib().orc_sumab(_p(out[0]), _p(out[1]), _p(out[2]), *[_p(f64(a)) for a in u], *[_p(f64(a)) for a in ulag1],
                    *[_p(f64(a)) for a in ulag2], abv, C.c_int(nab), C.c_int64(out[0].size))
    return out


def makeabf(ab1, ab2, f, rho, ext):
    """returns (ab1, ab2, f) after the update (copies)."""
    ab1, ab2, f = [f64(a).copy() for a in ab1], [f64(a).copy() for a in ab2], [f64(a).copy() for a in f]
    e = (C.c_double * 3)(*[float(a) for a in ext])
    lib().orc_makeabf(*[_p(a) for a in ab1], *[_p(a) for a in ab2], *[_p(a) for a in f], C.c_double(rho), e,
                      C.c_int64(f[0].size))
    return ab1, ab2, f


def makebdf(ulag1, ulag2, f, u, B, rho, dt, bd, nbd):
    f = [f64(a).copy() for a in f]
    b = (C.c_double * 4)(*[float(a) for a in bd])
    lib().orc_makebdf(*[_p(f64(a)) for a in ulag1], *[_p(f64(a)) for a in ulag2], *[_p(a) for a in f],
                      *[_p(f64(a)) for a in u], _p(f64(B)), C.c_double(rho), C.c_double(dt), b, C.c_int(nbd),
                      C.c_int64(f[0].size))
    return f


# ---- gather-scatter ----
def gs_classes(key):
    key = np.ascontiguousarray(key, dtype=np.int64)
    cid = np.zeros(key.size, dtype=np.int64)
    nc = lib().orc_gs_classes(cid.ctypes.data_as(_lp), key.ctypes.data_as(_lp), C.c_int64(key.size))
    return cid, int(nc)


def gs_add(f, cid, nclass):
    f = f64(f).copy()
    lib().orc_gs_add(_p(f), cid.ctypes.data_as(_lp), C.c_int64(nclass), C.c_int64(f.size))
    return f


class RhsStep:
    """Pre-bound arguments for repeated timing of one oracle "step" = orc_adjoint_rhs (source terms,
    mass matrix, adjoint advection, sensitivity) + orc_gs_add on the three components, i.e. the CPU
    restatement of adjoint_pnpn.f90:669-682 + :755-757.  Used by bench.py's cpu_baseline and
    --impl reference legs only (the one place outside tests/ that may execute oracle/)."""

    def __init__(self, v, vb, rho, lx, nelv, D, w, G, B, key, **kw):
        self.lx, self.nelv, self.n = lx, nelv, lx ** 3 * nelv
        self.a = dict(v=[f64(x) for x in v], vb=[f64(x) for x in vb], rho=f64(rho), D=_colmajor(D), w=f64(w),
                      G=[f64(g) for g in G], B=f64(B))
        self.f = [np.zeros(self.n) for _ in range(3)]
        self.sens = np.zeros(self.n)
        self.chi = np.zeros(self.n)
        self.p = params(kw.get("f_min", 0.0), kw.get("f_max", 1000.0), kw.get("q", 1.0), kw.get("convex_up", 1),
                        kw.get("if_lube", 1), kw.get("K_lube", 1.0), kw.get("K_sens", 1.0), kw.get("lxd", 0))
        self.cid, self.nc = gs_classes(key)
        self.Gp = _G(self.a["G"])

    def step(self):
        a, L = self.a, lib()
        L.orc_adjoint_rhs(_p(self.f[0]), _p(self.f[1]), _p(self.f[2]), _p(self.sens), _p(self.chi),
                          _p(a["v"][0]), _p(a["v"][1]), _p(a["v"][2]), _p(a["vb"][0]), _p(a["vb"][1]),
                          _p(a["vb"][2]), _p(a["rho"]), None, None, None, None, None, C.c_int(0),
                          C.c_int(self.lx), C.c_int(self.nelv), _p(a["D"]), _p(a["w"]), self.Gp, _p(a["B"]),
                          C.byref(self.p))
        for c in range(3):
            L.orc_gs_add(_p(self.f[c]), self.cid.ctypes.data_as(_lp), C.c_int64(self.nc), C.c_int64(self.n))
