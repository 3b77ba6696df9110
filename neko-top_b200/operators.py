"""Host-side mirror of the reference plug-in interface for the adjoint-RHS path, over the C ABI.

Names, argument meaning and error behaviour follow the reference's Fortran types so that the parity
tests read like tests of the reference (citations relative to /root/reference/sources):

  advection_adjoint_t / adv_lin_b200_t / advection_adjoint_factory   adjoint/advection_adjoint.f90:43-82,
                                                                      adjoint/advection_adjoint_fctry.f90:57-96
  simple_brinkman_source_term_t                                       source_terms/simple_brinkman_source_term.f90:52-153
  adjoint_lube_source_term_t                                          source_terms/adjoint_lube_source_term.f90:173-206
  RAMP_mapping_t                                                      mapping_functions/RAMP_mapping.f90:137-267
  gs_t (op with GS_OP_ADD)                                            adjoint/adjoint_pnpn.f90:725,755-757
  steady_simcomp_t                                                    simulation_components/steady_simcomp.f90:49-192
  fused_adjoint_rhs_t  -- the B200 path: adjoint_pnpn.f90:669-682 + :755-757 in one call

The Fortran text a maintainer adds to Neko-TOP is in fortran/ (INTEGRATION.md).  This Python layer exists
because pytest / bench.py drive the library; tensors are torch CUDA float64 tensors used purely as
device-memory handles.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check

GS_OP_ADD = 1


def _ptr(t):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype in (torch.float64, torch.int32, torch.int64) and t.is_contiguous()):
        raise ValueError("expected a contiguous CUDA tensor (float64 / int32 / int64)")
    return C.c_void_p(t.data_ptr())


def _ci(v):
    return C.byref(C.c_int(int(v)))


def _cd(v):
    return C.byref(C.c_double(float(v)))


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class space_t:
    """Subset of Neko's space_t: lx, dx (D(i,j)), wx."""

    def __init__(self, lx, dx, wx):
        self.lx, self.dx, self.wx = int(lx), np.asarray(dx, dtype=np.float64), np.asarray(wx, dtype=np.float64)
        self.lxyz = self.lx ** 3


class coef_t:
    """Subset of Neko's coef_t: device geometric factors (cofactors), B, optional jacinv; Xh; nelv."""

    def __init__(self, Xh, nelv, G, B, jacinv=None):
        self.Xh, self.nelv, self.G, self.B, self.jacinv = Xh, int(nelv), list(G), B, jacinv


class _handle:
    """Owns one b200 handle (one per coef_t), shared by the operator objects built on it."""

    def __init__(self, coef, device=None, stream=None):
        L = _lib.lib()
        self.coef = coef
        dev = torch.cuda.current_device() if device is None else device
        self.h = C.c_void_p()
        check(L.b200_adjrhs_create(C.byref(self.h), _ci(coef.Xh.lx), _ci(coef.nelv), _ci(dev)))
        dx = np.ascontiguousarray(coef.Xh.dx.T).reshape(-1)      # column-major D(i,j)
        wx = np.ascontiguousarray(coef.Xh.wx)
        check(L.b200_adjrhs_set_space(self.h, dx.ctypes.data_as(C.POINTER(C.c_double)),
                                      wx.ctypes.data_as(C.POINTER(C.c_double))))
        check(L.b200_adjrhs_set_stream(self.h, _stream_ptr(stream)))
        check(L.b200_adjrhs_set_geometry(self.h, *[_ptr(g) for g in coef.G], _ptr(coef.B)))
        self.n = coef.nelv * coef.Xh.lx ** 3

    def dealias_init(self, lxd=None):
        """adv_lin_dealias_t%init (adjoint/adv_adjoint_dealias.f90:137-161): hand the fine GL space and
        the GLL_to_GL matrix to the library, which interpolates the geometric factors (coef_GL)."""
        from . import sem
        if getattr(self, "lxd", 0) and (lxd is None or lxd == self.lxd):
            return
        ds = sem.DealiasSpace(self.coef.Xh.lx, lxd)
        dp = C.POINTER(C.c_double)
        J, Dd, wd = ds.interp_colmajor, ds.dxd_colmajor, np.ascontiguousarray(ds.wd)
        check(_lib.lib().b200_adv_dealias_init(self.h, _ci(ds.lxd), J.ctypes.data_as(dp), Dd.ctypes.data_as(dp),
                                               wd.ctypes.data_as(dp)))
        self.lxd = ds.lxd

    def free(self):
        if self.h:
            check(_lib.lib().b200_adjrhs_free(C.byref(self.h)))
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- advection_adjoint_t -----------------------------------------------------------------------------
class advection_adjoint_t:
    """adjoint/advection_adjoint.f90:43-50 (abstract)."""

    def compute_linear(self, vx, vy, vz, vxb, vyb, vzb, fx, fy, fz, Xh, coef, n):
        raise NotImplementedError

    def compute_adjoint(self, vx, vy, vz, vxb, vyb, vzb, fx, fy, fz, Xh, coef, n):
        raise NotImplementedError

    def free(self):
        raise NotImplementedError


class adv_lin_b200_t(advection_adjoint_t):
    """B200 replacement of adv_lin_no_dealias_t (adjoint/adv_adjoint_no_dealias.f90:57-77)."""

    def __init__(self):
        self._hd = None

    def init(self, coef, handle=None):
        self._hd = handle if handle is not None else _handle(coef)

    def compute_adjoint(self, vx, vy, vz, vxb, vyb, vzb, fx, fy, fz, Xh=None, coef=None, n=None):
        """f_i is IN/OUT: f_i -= (grad U_b)^T v |_i + weak-form div term (:119-255)."""
        if n is not None and n != self._hd.n:
            raise ValueError(f"n={n} does not match the handle ({self._hd.n})")
        check(_lib.lib().b200_adv_adjoint_compute(self._hd.h, _ptr(vx), _ptr(vy), _ptr(vz), _ptr(vxb),
                                                  _ptr(vyb), _ptr(vzb), _ptr(fx), _ptr(fy), _ptr(fz)))

    def compute_linear(self, vx, vy, vz, vxb, vyb, vzb, fx, fy, fz, Xh=None, coef=None, n=None):
        check(_lib.lib().b200_adv_linear_compute(self._hd.h, _ptr(vx), _ptr(vy), _ptr(vz), _ptr(vxb),
                                                 _ptr(vyb), _ptr(vzb), _ptr(self._hd.coef.jacinv),
                                                 _ptr(fx), _ptr(fy), _ptr(fz)))

    def free(self):
        if self._hd is not None:
            self._hd.free()
            self._hd = None


class adv_lin_dealias_b200_t(advection_adjoint_t):
    """B200 replacement of adv_lin_dealias_t (adjoint/adv_adjoint_dealias.f90:56-131)."""

    def __init__(self):
        self._hd = None

    def init(self, lxd, coef, handle=None):
        self._hd = handle if handle is not None else _handle(coef)
        self._hd.dealias_init(lxd)

    def compute_adjoint(self, vx, vy, vz, vxb, vyb, vzb, fx, fy, fz, Xh=None, coef=None, n=None):
        """adjoint/adv_adjoint_dealias.f90:235-462; f_i IN/OUT."""
        if n is not None and n != self._hd.n:
            raise ValueError(f"n={n} does not match the handle ({self._hd.n})")
        check(_lib.lib().b200_adv_adjoint_dealias_compute(self._hd.h, _ptr(vx), _ptr(vy), _ptr(vz), _ptr(vxb),
                                                          _ptr(vyb), _ptr(vzb), _ptr(fx), _ptr(fy), _ptr(fz)))

    def compute_linear(self, vx, vy, vz, vxb, vyb, vzb, fx, fy, fz, Xh=None, coef=None, n=None):
        """adjoint/adv_adjoint_dealias.f90:479-668; f_i IN/OUT."""
        check(_lib.lib().b200_adv_linear_dealias_compute(self._hd.h, _ptr(vx), _ptr(vy), _ptr(vz), _ptr(vxb),
                                                         _ptr(vyb), _ptr(vzb), _ptr(fx), _ptr(fy), _ptr(fz)))

    def free(self):
        if self._hd is not None:
            self._hd.free()
            self._hd = None


def advection_adjoint_factory(json, coef, handle=None):
    """adjoint/advection_adjoint_fctry.f90:57-96.  `json` is a dict with the case keys
    case.numerics.{dealias, polynomial_order, dealiased_polynomial_order}; one extra key,
    case.numerics.adjoint_backend = "b200", selects this implementation."""
    num = json.get("case", {}).get("numerics", {})
    dealias = bool(num.get("dealias", False))
    if dealias:
        lxd = num.get("dealiased_polynomial_order")
        if lxd is None:
            lxd = 3 * (int(num["polynomial_order"]) + 1) // 2       # :67-71 "assumes odd polynomial order"
        if lxd <= 0:
            lxd = coef.Xh.lx * 3 // 2                               # :89
        obj = adv_lin_dealias_b200_t()
        obj.init(lxd, coef, handle)
        return obj
    obj = adv_lin_b200_t()
    obj.init(coef, handle)
    return obj


# ---- source terms ----------------------------------------------------------------------------------------
class simple_brinkman_source_term_t:
    """source_terms/simple_brinkman_source_term.f90:52-153."""

    def init_from_components(self, f_x, f_y, f_z, design_chi, u, v, w, coef):
        self.fields = (f_x, f_y, f_z)
        self.u, self.v, self.w, self.chi = u, v, w, design_chi
        self.start_time, self.end_time = 0.0, 100000000.0      # :108-109

    def compute_(self, t=0.0, tstep=0):
        fu, fv, fw = self.fields
        check(_lib.lib().b200_brinkman_compute(_ptr(fu), _ptr(fv), _ptr(fw), _ptr(self.u), _ptr(self.v),
                                               _ptr(self.w), _ptr(self.chi), _ci(fu.numel()), _stream_ptr()))

    def free(self):
        self.fields = None


class adjoint_lube_source_term_t:
    """source_terms/adjoint_lube_source_term.f90:173-206; mask = 1-based int32 indices or None."""

    def init_from_components(self, f_x, f_y, f_z, design_chi, K, u, v, w, mask=None, if_mask=False, coef=None):
        self.fields = (f_x, f_y, f_z)
        self.u, self.v, self.w, self.chi, self.K = u, v, w, design_chi, float(K)
        self.mask = mask if if_mask else None

    def compute_(self, t=0.0, tstep=0):
        fu, fv, fw = self.fields
        ms = 0 if self.mask is None else self.mask.numel()
        check(_lib.lib().b200_lube_compute(_ptr(fu), _ptr(fv), _ptr(fw), _ptr(self.u), _ptr(self.v),
                                           _ptr(self.w), _ptr(self.chi), _cd(self.K), _ptr(self.mask),
                                           _ci(ms), _ci(fu.numel()), _stream_ptr()))


def opcolv(fx, fy, fz, B):
    """adjoint/adjoint_pnpn.f90:672-676."""
    check(_lib.lib().b200_opcolv(_ptr(fx), _ptr(fy), _ptr(fz), _ptr(B), _ci(fx.numel()), _stream_ptr()))


class RAMP_mapping_t:
    """mapping_functions/RAMP_mapping.f90 (defaults :107-110)."""

    def __init__(self, f_min=0.0, f_max=1000.0, q=1.0, convex_up=True):
        self.f_min, self.f_max, self.q, self.convex_up = f_min, f_max, q, convex_up

    def apply_forward(self, X_out, X_in):
        check(_lib.lib().b200_ramp_forward(_ptr(X_out), _ptr(X_in), _ci(X_in.numel()), _cd(self.f_min),
                                           _cd(self.f_max), _cd(self.q), _ci(self.convex_up), _stream_ptr()))

    def apply_backward(self, dF_dX_in, dF_dX_out, X_in):
        check(_lib.lib().b200_ramp_backward(_ptr(dF_dX_in), _ptr(dF_dX_out), _ptr(X_in), _ci(X_in.numel()),
                                            _cd(self.f_min), _cd(self.f_max), _cd(self.q),
                                            _ci(self.convex_up), _stream_ptr()))


def compute_sensitivity(sens, u, v, w, u_adj, v_adj, w_adj, K_obj=1.0, if_lube=True):
    """objectives/minimum_dissipation_objective_function.f90:260-301."""
    check(_lib.lib().b200_sensitivity(_ptr(sens), _ptr(u), _ptr(v), _ptr(w), _ptr(u_adj), _ptr(v_adj),
                                      _ptr(w_adj), _cd(K_obj), _ci(if_lube), _ci(sens.numel()), _stream_ptr()))


# ---- minimum-dissipation objective chain -------------------------------------------------------------------
def curl(handle, w, u, jacinv, Binv):
    """Neko curl(w1,w2,w3, u1,u2,u3, ., ., coef): strong curl, x B, gs_op(ADD), x Binv.  handle: a
    fused_adjoint_rhs_t (its gs must be initialised)."""
    check(_lib.lib().b200_curl(handle.handle.h, *[_ptr(a) for a in w], *[_ptr(a) for a in u], _ptr(jacinv), _ptr(Binv)))


class adjoint_minimum_dissipation_source_term_t:
    """source_terms/adjoint_minimum_dissipation_source_term.f90:52-249: f += obj_scale * curl(curl(u))."""

    def init_from_components(self, f_x, f_y, f_z, u, v, w, obj_scale, mask, if_mask, coef, handle, Binv):
        self.fields, self.u, self.v, self.w = (f_x, f_y, f_z), u, v, w
        self.obj_scale, self.mask, self.coef, self.handle, self.Binv = float(obj_scale), (mask if if_mask else None), coef, handle, Binv

    def compute_(self, t=0.0, tstep=0):
        ms = 0 if self.mask is None else self.mask.numel()
        check(_lib.lib().b200_curlcurl_forcing(self.handle.handle.h, *[_ptr(a) for a in self.fields], _ptr(self.u),
                                               _ptr(self.v), _ptr(self.w), _ptr(self.coef.jacinv), _ptr(self.Binv),
                                               _ptr(self.mask), _ci(ms), _cd(self.obj_scale)))


def min_dissipation_objective(handle, u, v, w, chi, jacinv, mask=None, K=1.0, obj_scale=1.0):
    """minimum_dissipation_objective_function_t%compute (:186-254), rank-local: (objective, dissipation, lube)."""
    out = (C.c_double * 3)()
    ms = 0 if mask is None else mask.numel()
    check(_lib.lib().b200_min_dissipation_objective(handle.handle.h, _ptr(u), _ptr(v), _ptr(w), _ptr(chi), _ptr(jacinv),
                                                    _ptr(mask), _ci(ms), _cd(K), _cd(obj_scale), out))
    return out[0], out[1], out[2]


def mask_exterior_const(fld, mask, const):
    """neko_ext/mask_ops.f90:55-82 on the device."""
    work = torch.empty_like(fld)
    check(_lib.lib().b200_mask_exterior_const(_ptr(fld), _ptr(work), _ptr(mask), _ci(mask.numel()), _cd(const),
                                              _ci(fld.numel()), _stream_ptr()))


class PDE_filter_t:
    """mapping_functions/PDE_filter_mapping.f90:52-363: apply_forward / apply_backward = one Helmholtz solve."""

    def __init__(self, handle, coef, mult, r=0.01, abs_tol=1e-10, max_iter=200, precond="ident", norm_fac=1.0,
                 x0_is_input=False):
        """Defaults = the reference's hard-coded attributes (PDE_filter_mapping.f90:131-137: r = 0.01,
        abstol 1e-10, 200 iterations, "ident" preconditioner).  x0_is_input: start the Krylov iteration from the
        unfiltered field (the intent of :246-248) instead of from zero (what Neko's solvers do on entry)."""
        self.handle, self.coef, self.mult = handle, coef, mult
        self.r, self.abs_tol, self.max_iter, self.precond, self.norm_fac = r, abs_tol, max_iter, precond, norm_fac
        self.x0_is_input = bool(x0_is_input)
        self.ksp_results = None

    def _solve(self, X_out, X_in):
        it, r0, r1 = C.c_int(0), C.c_double(0), C.c_double(0)
        check(_lib.lib().b200_pde_filter_apply(self.handle.handle.h, _ptr(X_out), _ptr(X_in), _ptr(self.coef.jacinv),
                                               _ptr(self.mult), _cd(self.r), _cd(self.abs_tol), _ci(self.max_iter),
                                               _ci(0 if self.precond == "ident" else 1), _cd(self.norm_fac),
                                               _ci(self.x0_is_input), C.byref(it), C.byref(r0), C.byref(r1)))
        self.ksp_results = (it.value, r0.value, r1.value)

    def apply_forward(self, X_out, X_in):
        self._solve(X_out, X_in)

    def apply_backward(self, dF_dX_in, dF_dX_out, X_in=None):
        self._solve(dF_dX_in, dF_dX_out)


# ---- explicit time scheme around the RHS (Neko rhs_maker types; adjoint_pnpn.f90:665-666,688-696) ---------
def _dv(a):
    v = np.ascontiguousarray(a, dtype=np.float64)
    return v, v.ctypes.data_as(C.POINTER(C.c_double))


def _p3(a):
    return [None] * 3 if a is None else [_ptr(t) for t in a]


class rhs_maker_sumab_t:
    """sumab%compute_fluid(u_e, v_e, w_e, u, v, w, ulag, vlag, wlag, ab, nab); *lag = (lag1[, lag2]) tensors."""

    def compute_fluid(self, u_e, v_e, w_e, u, v, w, ulag, vlag, wlag, ab, nab):
        _, abp = _dv(list(ab) + [0.0] * (3 - len(ab)))
        l1 = [ulag[0], vlag[0], wlag[0]]
        l2 = [ulag[1], vlag[1], wlag[1]] if nab == 3 else None
        check(_lib.lib().b200_sumab(_ptr(u_e), _ptr(v_e), _ptr(w_e), _ptr(u), _ptr(v), _ptr(w), *_p3(l1), *_p3(l2),
                                    abp, _ci(nab), _ci(u.numel()), _stream_ptr()))


class rhs_maker_ext_t:
    """makeabf%compute_fluid(abx1, aby1, abz1, abx2, aby2, abz2, f_x, f_y, f_z, rho, ext, n)."""

    def compute_fluid(self, abx1, aby1, abz1, abx2, aby2, abz2, fx, fy, fz, rho, ext, n=None):
        _, ep = _dv(ext)
        check(_lib.lib().b200_makeabf(_ptr(abx1), _ptr(aby1), _ptr(abz1), _ptr(abx2), _ptr(aby2), _ptr(abz2),
                                      _ptr(fx), _ptr(fy), _ptr(fz), _cd(rho), ep, _ci(fx.numel()), _stream_ptr()))


class rhs_maker_bdf_t:
    """makebdf%compute_fluid(ulag, vlag, wlag, f_x, f_y, f_z, u, v, w, B, rho, dt, bd, nbd, n)."""

    def compute_fluid(self, ulag, vlag, wlag, fx, fy, fz, u, v, w, B, rho, dt, bd, nbd, n=None):
        _, bp = _dv(list(bd) + [0.0] * (4 - len(bd)))
        l1 = [ulag[0], vlag[0], wlag[0]] if nbd >= 2 else None
        l2 = [ulag[1], vlag[1], wlag[1]] if nbd >= 3 else None
        check(_lib.lib().b200_makebdf(*_p3(l1), *_p3(l2), _ptr(fx), _ptr(fy), _ptr(fz), _ptr(u), _ptr(v), _ptr(w),
                                      _ptr(B), _cd(rho), _cd(dt), bp, _ci(nbd), _ci(fx.numel()), _stream_ptr()))


def makeabf_bdf(ab1, ab2, ulag, vlag, wlag, f, u, B, rho, dt, ext, bd, nbd):
    """makeabf + makebdf in one pass over f (b200_makeabf_bdf)."""
    _, ep = _dv(ext)
    _, bp = _dv(list(bd) + [0.0] * (4 - len(bd)))
    l1 = [ulag[0], vlag[0], wlag[0]] if nbd >= 2 else None
    l2 = [ulag[1], vlag[1], wlag[1]] if nbd >= 3 else None
    check(_lib.lib().b200_makeabf_bdf(*_p3(ab1), *_p3(ab2), *_p3(l1), *_p3(l2), *_p3(f), *_p3(u), _ptr(B), _cd(rho),
                                      _cd(dt), ep, bp, _ci(nbd), _ci(f[0].numel()), _stream_ptr()))


class steady_simcomp_t:
    """simulation_components/steady_simcomp.f90:49-192 for a list of device fields."""

    def init_from_attributes(self, tol, fields):
        self.tol = float(tol)
        self.fields = list(fields)
        self.old = [torch.zeros_like(f) for f in fields]
        self.freeze = False

    def compute_(self, t=0.0, tstep=0):
        if self.freeze:
            return
        normed = []
        for f, fo in zip(self.fields, self.old):
            r = C.c_double(0.0)
            check(_lib.lib().b200_steady_field_update(C.byref(r), _ptr(f), _ptr(fo), _ci(f.numel()), _stream_ptr()))
            normed.append(r.value)
        self.normed_diff = normed
        if max(normed) <= self.tol:
            self.freeze = True


# ---- gather-scatter ----------------------------------------------------------------------------------------
class gs_t:
    """Neko gs_t restricted to op(., GS_OP_ADD) (adjoint/adjoint_scheme.f90:339-343 builds it from the
    dofmap; here from the global node keys)."""

    def __init__(self, handle):
        self._hd = handle

    def init(self, keys):
        """keys: int64 tensor (n) on the handle's device, or a host numpy array."""
        if isinstance(keys, torch.Tensor) and keys.is_cuda:
            k = keys.contiguous().view(-1)
            assert k.dtype == torch.int64 and k.numel() == self._hd.n
            check(_lib.lib().b200_gs_init(self._hd.h, _ptr(k), _ci(1)))
            torch.cuda.synchronize()
        else:
            k = np.ascontiguousarray(np.asarray(keys).reshape(-1), dtype=np.int64)
            assert k.size == self._hd.n
            check(_lib.lib().b200_gs_init(self._hd.h, k.ctypes.data_as(C.POINTER(C.c_int64)), _ci(0)))

    def classes(self):
        """(class_id[n] int64 numpy, nclass): canonical labelling (parity hook)."""
        cid = np.zeros(self._hd.n, dtype=np.int64)
        nc = C.c_int64(0)
        check(_lib.lib().b200_gs_get_classes(self._hd.h, cid.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(nc)))
        return cid, nc.value

    def op(self, f, op=GS_OP_ADD):
        if op != GS_OP_ADD:
            raise ValueError("only GS_OP_ADD is on the adjoint-RHS path")
        check(_lib.lib().b200_gs_op(self._hd.h, _ptr(f)))

    def op3(self, fx, fy, fz):
        check(_lib.lib().b200_gs_op3(self._hd.h, _ptr(fx), _ptr(fy), _ptr(fz)))

    def init_shared(self, shared_dof, neigh_rank, neigh_off, neigh_idx):
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (shared_dof, neigh_rank, neigh_off, neigh_idx)]
        ip = C.POINTER(C.c_int)
        check(_lib.lib().b200_gs_init_shared(self._hd.h, _ci(a[0].size), a[0].ctypes.data_as(ip),
                                             _ci(a[1].size), a[1].ctypes.data_as(ip), a[2].ctypes.data_as(ip),
                                             a[3].ctypes.data_as(ip)))

    def init_shared_from_keys(self, keys, candidates=None):
        """Shared-node discovery behind the C ABI (b200_gs_init_shared_from_keys; collective over the
        communicator): keys as in init(), candidates = optional uint8/bool mask of the dofs that can live on
        another rank (Neko: dm_Xh%shared_dof).  Returns (nshared, nneigh)."""
        ns, nn = C.c_int(0), C.c_int(0)
        if isinstance(keys, torch.Tensor) and keys.is_cuda:
            k = keys.contiguous().view(-1)
            assert k.dtype == torch.int64 and k.numel() == self._hd.n
            c = None
            if candidates is not None:
                c = candidates.contiguous().view(-1).to(torch.uint8)
                assert c.is_cuda and c.numel() == self._hd.n
            cp = None if c is None else C.c_void_p(c.data_ptr())
            check(_lib.lib().b200_gs_init_shared_from_keys(self._hd.h, _ptr(k), _ci(1), cp, C.byref(ns), C.byref(nn)))
            torch.cuda.synchronize()
        else:
            k = np.ascontiguousarray(np.asarray(keys).reshape(-1), dtype=np.int64)
            cp = None
            if candidates is not None:
                c = np.ascontiguousarray(np.asarray(candidates).reshape(-1), dtype=np.uint8)
                cp = c.ctypes.data_as(C.c_void_p)
            check(_lib.lib().b200_gs_init_shared_from_keys(self._hd.h, k.ctypes.data_as(C.POINTER(C.c_int64)), _ci(0),
                                                           cp, C.byref(ns), C.byref(nn)))
        return ns.value, nn.value


# ---- the fused B200 path -----------------------------------------------------------------------------------
class fused_adjoint_rhs_t:
    """One object per coef_t.  compute(): source terms + mass matrix + adjoint advection + sensitivity
    in one kernel pass (adjoint_pnpn.f90:669-682); step(): compute() + gs_op on f (:755-757)."""

    def __init__(self, coef, device=None, stream=None):
        self._hd = _handle(coef, device, stream)
        self.gs = gs_t(self._hd)
        self.n = self._hd.n

    @property
    def handle(self):
        return self._hd

    def set_params(self, f_min=0.0, f_max=1000.0, q=1.0, convex_up=True, if_lube=True, K_lube=1.0, K_sens=1.0):
        check(_lib.lib().b200_adjrhs_set_params(self._hd.h, _cd(f_min), _cd(f_max), _cd(q), _ci(convex_up),
                                                _ci(if_lube), _cd(K_lube), _cd(K_sens)))

    def set_dealias(self, flag=True, lxd=None):
        """compute()/step() use the dealiased adjoint operator (case.numerics.dealias = true)."""
        if flag:
            self._hd.dealias_init(lxd)
        check(_lib.lib().b200_adjrhs_set_dealias(self._hd.h, _ci(bool(flag))))

    def set_lube_mask(self, mask):
        self._mask = mask
        check(_lib.lib().b200_adjrhs_set_lube_mask(self._hd.h, _ptr(mask), _ci(0 if mask is None else mask.numel())))

    def _args(self, v, vb, rho, chi, fstatic, f, sens, chi_out):
        fs = fstatic if fstatic is not None else (None, None, None)
        return ([_ptr(a) for a in v] + [_ptr(a) for a in vb] + [_ptr(rho), _ptr(chi)]
                + [_ptr(a) for a in fs] + [_ptr(a) for a in f] + [_ptr(sens), _ptr(chi_out)])

    def compute(self, v, vb, f, rho=None, chi=None, fstatic=None, sens=None, chi_out=None):
        check(_lib.lib().b200_adjrhs_compute(self._hd.h, *self._args(v, vb, rho, chi, fstatic, f, sens, chi_out)))

    def step(self, v, vb, f, rho=None, chi=None, fstatic=None, sens=None, chi_out=None):
        check(_lib.lib().b200_adjrhs_step(self._hd.h, *self._args(v, vb, rho, chi, fstatic, f, sens, chi_out)))

    def step_host(self, v, vb, rho, f, sens=None):
        """HOST (pinned) float64 torch/numpy buffers in and out; bench.py's e2e."""
        def hp(a):
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                assert (not a.is_cuda) and a.dtype == torch.float64 and a.is_contiguous()
                return C.c_void_p(a.data_ptr())
            return a.ctypes.data_as(C.c_void_p)
        check(_lib.lib().b200_adjrhs_step_host(self._hd.h, *[hp(a) for a in v], *[hp(a) for a in vb], hp(rho),
                                               *[hp(a) for a in f], hp(sens)))

    def comm_init(self, id_bytes, rank, nranks):
        buf = C.create_string_buffer(bytes(id_bytes), 128)
        check(_lib.lib().b200_comm_init(self._hd.h, buf, _ci(rank), _ci(nranks)))

    def set_boundary_elements(self, elems):
        e = np.ascontiguousarray(elems, dtype=np.int32)
        check(_lib.lib().b200_adjrhs_set_boundary_elements(self._hd.h, _ci(e.size), e.ctypes.data_as(C.POINTER(C.c_int))))

    def set_element_order(self, order):
        """Processing order of the elements (permutation of 0..nelv-1; None: mesh order)."""
        if order is None:
            check(_lib.lib().b200_adjrhs_set_element_order(self._hd.h, _ci(0), None))
            return
        o = np.ascontiguousarray(order, dtype=np.int32)
        check(_lib.lib().b200_adjrhs_set_element_order(self._hd.h, _ci(o.size), o.ctypes.data_as(C.POINTER(C.c_int))))

    def set_gs_mode(self, mode):
        """Class-list pass when the staged summation is not in use: 0 = CSR lists (default), 1 = lists packed by size."""
        check(_lib.lib().b200_adjrhs_set_gs_fused(self._hd.h, _ci(1 if int(mode) else 0)))

    def gs_info(self):
        """(0, classes in the packed lists when in use, classes in total)."""
        a, b, c = C.c_int(0), C.c_int64(0), C.c_int64(0)
        check(_lib.lib().b200_adjrhs_gs_info(self._hd.h, C.byref(a), C.byref(b), C.byref(c)))
        return bool(a.value), b.value, c.value

    def set_xstage(self, level=2):
        """Staged direct-stiffness summation at lx = 8: 0 off, 1 = x pairs in the element kernel (bit-identical),
        2 = product classes summed direction by direction (x in the kernel, y / z face passes; default)."""
        check(_lib.lib().b200_adjrhs_set_xstage(self._hd.h, _ci(int(level))))

    def xstage_info(self):
        """(level in use, classes staged, classes left to the class-list pass, classes in total)."""
        a, b, c, d = C.c_int(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(_lib.lib().b200_adjrhs_xstage_info(self._hd.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    def enable_timing(self, flag=True):
        check(_lib.lib().b200_adjrhs_enable_timing(self._hd.h, _ci(flag)))

    def get_timing(self):
        a, b, c = C.c_double(0), C.c_double(0), C.c_int64(0)
        check(_lib.lib().b200_adjrhs_get_timing(self._hd.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def get_phase_timing(self):
        ms = (C.c_double * 10)()
        n = C.c_int(10)
        check(_lib.lib().b200_adjrhs_get_phase_timing(self._hd.h, ms, C.byref(n)))
        return [ms[i] for i in range(n.value)]

    def free(self):
        self._hd.free()


def comm_unique_id():
    buf = C.create_string_buffer(128)
    check(_lib.lib().b200_comm_unique_id(buf))
    return bytes(buf.raw)


def launch_count():
    return int(_lib.lib().b200_launch_count())
