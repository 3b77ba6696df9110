"""numpy twin of oracle/oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

An independent restatement (numpy.polynomial for the quadrature, einsum for the tensor
contractions, whole-field "device branch" ordering) used to cross-check the C oracle on small
cases.  Citations are relative to /root/reference.  Parity unpinned: see oracle/oracle.h.

Arrays are handled as (nelv, lz, ly, lx) C-order views == Fortran x(lx,ly,lz,nelv).
"""
import numpy as np
from numpy.polynomial import legendre as L


def gll(n):
    N = n - 1
    PN = L.Legendre.basis(N)
    z = np.concatenate(([-1.0], np.sort(PN.deriv().roots().real), [1.0])) if n > 2 else np.array([-1.0, 1.0])
    z = 0.5 * (z - z[::-1])
    w = 2.0 / (N * (N + 1) * PN(z) ** 2)
    return z, w


def gl(n):
    z, w = L.leggauss(n)
    return z, w


def lagrange_deriv(z):
    n = len(z)
    bw = np.array([1.0 / np.prod([z[j] - z[k] for k in range(n) if k != j]) for j in range(n)])
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[i, j] = (bw[j] / bw[i]) / (z[i] - z[j])
        D[i, i] = -D[i].sum()
    return D


def dgll(z):
    n = len(z)
    N = n - 1
    PN = L.Legendre.basis(N)(z)
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[i, j] = PN[i] / (PN[j] * (z[i] - z[j]))
    D[0, 0] = -N * (N + 1) / 4.0
    D[N, N] = N * (N + 1) / 4.0
    return D


def interp(zto, zfrom):
    n = len(zfrom)
    J = np.zeros((len(zto), n))
    for m in range(n):
        c = np.zeros(n)
        c[m] = 1.0
        # l_m(x) = prod_{k != m} (x - z_k)/(z_m - z_k)
        num = np.ones(len(zto))
        for k in range(n):
            if k != m:
                num *= (zto - zfrom[k]) / (zfrom[m] - zfrom[k])
        J[:, m] = num
    return J


def v4(a, lx, nelv):
    return np.asarray(a, dtype=np.float64).reshape(nelv, lx, lx, lx)   # [e,k,j,i]


def dr(u, D):
    return np.einsum("im,ekjm->ekji", D, u)


def ds(u, D):
    return np.einsum("jm,ekmi->ekji", D, u)


def dt(u, D):
    return np.einsum("km,emji->ekji", D, u)


def drT(u, D):
    return np.einsum("mi,ekjm->ekji", D, u)


def dsT(u, D):
    return np.einsum("mj,ekmi->ekji", D, u)


def dtT(u, D):
    return np.einsum("mk,emji->ekji", D, u)


def w3_of(w):
    return np.einsum("k,j,i->kji", w, w, w)


def geom(x, y, z, D, w):
    """cofactors (J-scaled), jac, B.  x,y,z: (nelv,lx,lx,lx)."""
    xr, xs, xt = dr(x, D), ds(x, D), dt(x, D)
    yr, ys, yt = dr(y, D), ds(y, D), dt(y, D)
    zr, zs, zt = dr(z, D), ds(z, D), dt(z, D)
    jac = (xr * ys * zt + xt * yr * zs + xs * yt * zr - xr * yt * zs - xs * yr * zt - xt * ys * zr)
    G = [ys * zt - yt * zs, yt * zr - yr * zt, yr * zs - ys * zr,
         xt * zs - xs * zt, xr * zt - xt * zr, xs * zr - xr * zs,
         xs * yt - xt * ys, xt * yr - xr * yt, xr * ys - xs * yr]
    return G, jac, jac * w3_of(w)[None]


def opgrad(u, D, w3, G):
    ur, us, ut = dr(u, D), ds(u, D), dt(u, D)
    return (w3 * (G[0] * ur + G[1] * us + G[2] * ut),
            w3 * (G[3] * ur + G[4] * us + G[5] * ut),
            w3 * (G[6] * ur + G[7] * us + G[8] * ut))


def cdtp(x, gr, gs_, gt, D, w3):
    wx = x * w3
    return drT(wx * gr, D) + dsT(wx * gs_, D) + dtT(wx * gt, D)


def adjoint_advection(f, v, vb, D, w, G):
    """adv_adjoint_no_dealias.f90:162-201 + :269-303 (device branch).  Lists of (nelv,l,l,l)."""
    w3 = w3_of(w)[None]
    f = [a.copy() for a in f]
    g = [opgrad(vb[c], D, w3, G) for c in range(3)]          # g[c][d] = d(U_c)/dx_d (weak)
    for d in range(3):
        f[d] -= v[0] * g[0][d] + v[1] * g[1][d] + v[2] * g[2][d]
    for c in range(3):
        acc = 0
        for k in range(3):
            acc = acc + cdtp(v[c] * vb[k], G[3 * k], G[3 * k + 1], G[3 * k + 2], D, w3)
        f[c] -= acc
    return f


def conv1(u, vx, vy, vz, D, G, jacinv):
    ur, us, ut = dr(u, D), ds(u, D), dt(u, D)
    return jacinv * (vx * (G[0] * ur + G[1] * us + G[2] * ut) + vy * (G[3] * ur + G[4] * us + G[5] * ut)
                     + vz * (G[6] * ur + G[7] * us + G[8] * ut))


def linear_advection(f, v, vb, D, w, G, jac):
    """adv_adjoint_no_dealias.f90:404-424"""
    B = jac * w3_of(w)[None]
    ji = 1.0 / jac
    f = [a.copy() for a in f]
    for c in range(3):
        f[c] -= B * conv1(v[c], vb[0], vb[1], vb[2], D, G, ji)
        f[c] -= B * conv1(vb[c], v[0], v[1], v[2], D, G, ji)
    return f


def tens(u, A):
    return np.einsum("cn,bm,al,enml->ecba", A, A, A, u)


def adjoint_advection_dealias(f, v, vb, lx, lxd, G):
    """adv_adjoint_dealias.f90:137-161, 264-351 (device-branch ordering, whole field)."""
    zg, _ = gll(lx)
    zd, wd = gl(lxd)
    J = interp(zd, zg)
    Dd = lagrange_deriv(zd)
    w3d = w3_of(wd)[None]
    Gd = [tens(g, J) for g in G]
    t = [tens(a, J) for a in v]
    tb = [tens(a, J) for a in vb]
    f = [a.copy() for a in f]
    g = [opgrad(tb[c], Dd, w3d, Gd) for c in range(3)]
    for d in range(3):
        f[d] -= tens(t[0] * g[0][d] + t[1] * g[1][d] + t[2] * g[2][d], J.T)
    for c in range(3):
        acc = 0
        for k in range(3):
            acc = acc + cdtp(t[c] * tb[k], Gd[3 * k], Gd[3 * k + 1], Gd[3 * k + 2], Dd, w3d)
        f[c] -= tens(acc, J.T)
    return f


def ramp(rho, f_min=0.0, f_max=1000.0, q=1.0, convex_up=True):
    if convex_up:
        return f_min + (f_max - f_min) * rho * (1.0 + q) / (rho + q)
    return f_min + (f_max - f_min) * rho / (1.0 + q * (1.0 - rho))


def adjoint_rhs(v, vb, D, w, G, B, chi, K_lube=1.0, if_lube=True, fstatic=None):
    """adjoint_pnpn.f90:661-682 with the steady-problem source terms."""
    f = [np.zeros_like(v[0]) for _ in range(3)]
    for c in range(3):
        f[c] -= v[c] * chi
        if fstatic is not None:
            f[c] += fstatic[c]
        if if_lube:
            f[c] += vb[c] * (chi * K_lube)
        f[c] *= B
    return adjoint_advection(f, v, vb, D, w, G)


def sensitivity(u, ua, K_obj=1.0, if_lube=True):
    s = -(u[0] * ua[0] + u[1] * ua[1] + u[2] * ua[2])
    if if_lube:
        s = s + K_obj * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2])
    return s


def gs_add(f, key):
    _, inv = np.unique(key, return_inverse=True)
    acc = np.zeros(inv.max() + 1)
    np.add.at(acc, inv, f)
    return acc[inv]


def steady_simcomp_compute(fields, olds, tol):
    """simulation_components/steady_simcomp.f90:154-188 restated: old -= new (field_sub2, :155-161),
    normed_diff = glsc2(old, old) per field (:165-173), then EITHER old = new (field_copy, :178-186) when
    max(normed_diff) > tol OR freeze (:188).  `olds` is updated in place like the reference's this%u_old ...;
    returns (normed_diff list, freeze).  Sums are left-to-right in float64 extended by math.fsum (exact), the
    CUDA path's deterministic tree differs from it only by rounding."""
    import math
    normed = []
    for f, o in zip(fields, olds):
        o -= f
        normed.append(math.fsum((o * o).tolist()))
    freeze = not (max(normed) > tol)
    if not freeze:
        for f, o in zip(fields, olds):
            o[:] = f
    return normed, freeze
