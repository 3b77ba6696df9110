"""What Neko's `space_t` and `coef_t` hand to the adjoint-RHS path, rebuilt on the host/GPU.

In a Neko-TOP run these arrays already exist (Xh%dx, Xh%wx, coef%drdx_d ... coef%B_d) and are passed
straight through the C ABI.  For the stand-alone workloads of BASELINE.json there is
no Neko, so this module produces the same quantities:

  * GLL / GL points and weights, the GLL derivative matrix (Nek5000 speclib zwgll / zwgl / dgll),
    Lagrange interpolation and derivative matrices (Neko `setup_intp`);
  * geometric factors in Neko's convention: cofactors drdx..dtdz NOT divided by the
    Jacobian, jac, and B = jac*w3.

torch is used only as the array library (CPU in tests, CUDA in bench.py); nothing here is on the
timed path.
"""
import math

import numpy as np
import torch


def _legendre(x, n):
    """P_n(x), P_n'(x) by recurrence (numpy, float64)."""
    x = np.asarray(x, dtype=np.float64)
    p0, p1 = np.ones_like(x), x.copy()
    d0, d1 = np.zeros_like(x), np.ones_like(x)
    if n == 0:
        return p0, d0
    for k in range(2, n + 1):
        pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
        dk = d0 + (2 * k - 1) * p1
        p0, p1, d0, d1 = p1, pk, d1, dk
    return p1, d1


def zwgll(n):
    """Gauss-Lobatto-Legendre nodes and weights on [-1,1] (n points)."""
    N = n - 1
    z = -np.cos(np.pi * np.arange(n) / N)
    for _ in range(100):
        p, dp = _legendre(z[1:-1], N)
        ddp = (2 * z[1:-1] * dp - N * (N + 1) * p) / (1 - z[1:-1] ** 2)
        dz = dp / ddp
        z[1:-1] -= dz
        if np.max(np.abs(dz), initial=0.0) < 1e-16:
            break
    z = 0.5 * (z - z[::-1])
    z[0], z[-1] = -1.0, 1.0
    p, _ = _legendre(z, N)
    return z, 2.0 / (N * (N + 1) * p * p)


def zwgl(n):
    """Gauss-Legendre nodes and weights (n points)."""
    z = -np.cos(np.pi * (np.arange(n) + 0.75) / (n + 0.5))
    for _ in range(100):
        p, dp = _legendre(z, n)
        dz = p / dp
        z -= dz
        if np.max(np.abs(dz)) < 1e-16:
            break
    z = 0.5 * (z - z[::-1])
    _, dp = _legendre(z, n)
    return z, 2.0 / ((1 - z * z) * dp * dp)


def dgll(z):
    """GLL derivative matrix D[i, j] = l_j'(z_i)."""
    n = len(z)
    N = n - 1
    p, _ = _legendre(z, N)
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[i, j] = p[i] / (p[j] * (z[i] - z[j]))
    D[0, 0] = -N * (N + 1) / 4.0
    D[N, N] = N * (N + 1) / 4.0
    return D


def _bary(z):
    n = len(z)
    return np.array([1.0 / np.prod([z[j] - z[k] for k in range(n) if k != j]) for j in range(n)])


def deriv_matrix(z):
    """Derivative matrix of the Lagrange basis on arbitrary nodes (GL-space dx in Neko)."""
    n = len(z)
    bw = _bary(z)
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[i, j] = (bw[j] / bw[i]) / (z[i] - z[j])
        D[i, i] = -np.sum(D[i])
    return D


def interp_matrix(zto, zfrom):
    """J[a, m] = l_m(zto_a) (Neko interpolator_t, GLL -> GL)."""
    bw = _bary(zfrom)
    J = np.zeros((len(zto), len(zfrom)))
    for a, x in enumerate(zto):
        d = x - zfrom
        hit = np.where(d == 0.0)[0]
        if len(hit):
            J[a, hit[0]] = 1.0
        else:
            t = bw / d
            J[a] = t / t.sum()
    return J


class Space:
    """Subset of Neko's space_t used by the path."""

    def __init__(self, lx):
        self.lx = lx
        self.zg, self.wx = zwgll(lx)
        self.dx = dgll(self.zg)                   # dx[i, j] = D(i, j)
        self.lxyz = lx ** 3

    @property
    def dx_colmajor(self):
        """Flat column-major buffer == Fortran Xh%dx(lx,lx) as the C ABI expects it."""
        return np.ascontiguousarray(self.dx.T).reshape(-1)

    def w3(self):
        return np.einsum("k,j,i->kji", self.wx, self.wx, self.wx)


class DealiasSpace:
    """What adv_lin_dealias_t%init builds (adjoint/adv_adjoint_dealias.f90:137-146): the fine
    Gauss-Legendre space Xh_GL (lxd points: nodes, weights, derivative matrix) and the GLL -> GL
    interpolation matrix of GLL_to_GL.  Default lxd = 3*lx/2 (advection_adjoint_fctry.f90:70,89)."""

    def __init__(self, lx, lxd=None):
        self.lx = lx
        self.lxd = 3 * lx // 2 if lxd is None else int(lxd)
        zg, _ = zwgll(lx)
        self.zd, self.wd = zwgl(self.lxd)
        self.interp = interp_matrix(self.zd, zg)        # [a, l] = J(a, l)
        self.dxd = deriv_matrix(self.zd)                # [i, j] = D(i, j)

    @property
    def interp_colmajor(self):
        return np.ascontiguousarray(self.interp.T).reshape(-1)

    @property
    def dxd_colmajor(self):
        return np.ascontiguousarray(self.dxd.T).reshape(-1)


def geometric_factors(x, y, z, space, chunk=32768):
    """coef_t arrays for nodal coordinates x,y,z of shape (nelv, lx, lx, lx) [e,k,j,i] (torch, any
    device).  Returns (G, jac, B): G = [drdx,dsdx,dtdx, drdy,dsdy,dtdy, drdz,dsdz,dtdz]."""
    dev = x.device
    D = torch.as_tensor(space.dx, dtype=torch.float64, device=dev)
    w3 = torch.as_tensor(space.w3(), dtype=torch.float64, device=dev)
    nelv = x.shape[0]
    G = [torch.empty_like(x) for _ in range(9)]
    jac = torch.empty_like(x)
    B = torch.empty_like(x)

    def dr(u):
        return torch.einsum("im,ekjm->ekji", D, u)

    def ds(u):
        return torch.einsum("jm,ekmi->ekji", D, u)

    def dt(u):
        return torch.einsum("km,emji->ekji", D, u)

    for e0 in range(0, nelv, chunk):
        s = slice(e0, min(nelv, e0 + chunk))
        xx, yy, zz = x[s], y[s], z[s]
        xr, xs, xt = dr(xx), ds(xx), dt(xx)
        yr, ys, yt = dr(yy), ds(yy), dt(yy)
        zr, zs, zt = dr(zz), ds(zz), dt(zz)
        J = (xr * ys * zt + xt * yr * zs + xs * yt * zr - xr * yt * zs - xs * yr * zt - xt * ys * zr)
        G[0][s] = ys * zt - yt * zs
        G[1][s] = yt * zr - yr * zt
        G[2][s] = yr * zs - ys * zr
        G[3][s] = xt * zs - xs * zt
        G[4][s] = xr * zt - xt * zr
        G[5][s] = xs * zr - xr * zs
        G[6][s] = xs * yt - xt * ys
        G[7][s] = xt * yr - xr * yt
        G[8][s] = xr * ys - xs * yr
        jac[s] = J
        B[s] = J * w3
    return G, jac, B


def phi_surface(lx):
    """Fraction of an element's points lying on its surface."""
    return 1.0 - ((lx - 2) / lx) ** 3


def algorithmic_bytes_per_dof(lx, with_gs=True):
    """SURVEY.md 8d contract: 168 B/DOF for the element kernel + 56*phi(lx) for gs_op on 3 fields."""
    return 168.0 + (56.0 * phi_surface(lx) if with_gs else 0.0)


def algorithmic_flops_per_dof(lx):
    return 36.0 * lx + 136.0


__all__ = ["zwgll", "zwgl", "dgll", "deriv_matrix", "interp_matrix", "Space", "DealiasSpace", "geometric_factors",
           "phi_surface", "algorithmic_bytes_per_dof", "algorithmic_flops_per_dof", "math"]
