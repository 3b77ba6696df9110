/*
 * oracle.h -- CPU restatement (plain C, fp64) of Neko-TOP's adjoint-RHS hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / the timed CPU baseline.  The product path (neko-top_b200/) never
 * calls into this library and has no CPU fallback.
 *
 * PARITY UNPINNED: the reference tree holds no golden vector, known-answer test or fixture for
 * this path, and the reference cannot be built here (no Fortran, no
 * MPI, Neko itself is an un-vendored, unpinned `develop` dependency: scripts/dependencies.sh:181-187).
 * The oracle is therefore anchored on (i) the reference's own call sites, restated line by line
 * below, (ii) the published Nek5000/Neko operator definitions (opgrad, cdtp, conv1, tnsr3d,
 * speclib zwgll/zwgl/dgll) and (iii) mathematical identities checked in tests/ (polynomial
 * exactness, discrete adjoint identity <v, L u> = <L^T v, u>).
 *
 * Memory layout everywhere: Fortran column-major x(lx,lx,lx,nelv), i fastest, element slowest.
 * All citations are relative to /root/reference.
 */
#ifndef NEKOTOP_ORACLE_H
#define NEKOTOP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- speclib restatements (Neko math/speclib.f90 = Nek5000 speclib; not vendored) ---------- */
void orc_zwgll(double *z, double *w, int np);              /* Gauss-Lobatto-Legendre nodes+weights */
void orc_zwgl(double *z, double *w, int np);               /* Gauss-Legendre nodes+weights        */
void orc_dgll(double *D, const double *z, int np);         /* D(i,j) col-major: D[i + np*j]       */
void orc_deriv_matrix(double *D, const double *z, int np); /* Lagrange derivative on any nodes    */
void orc_interp_matrix(double *J, const double *zto, int nto, const double *zfrom, int nfrom);
                                                           /* J(a,m) col-major: J[a + nto*m]      */

/* ---- coef_t restatement: geometric factors from nodal coordinates (Neko sem/coef.f90) -------
 * G[0..8] = drdx,dsdx,dtdx, drdy,dsdy,dtdy, drdz,dsdz,dtdz  (cofactors, NOT divided by jac),
 * jac, B = jac*w3.  SURVEY.md 8c lists the formulas. */
void orc_geom(int lx, int nelv, const double *D, const double *w,
              const double *x, const double *y, const double *z,
              double *G[9], double *jac, double *B);

/* ---- Neko operators the path calls ------------------------------------------ */
/* tensor-product interpolation v = (A x A x A) u, A is (nv x nu) col-major; per element */
void orc_tnsr3d(double *v, int nv, const double *u, int nu, const double *A, int nelv);
/* transpose map: u = (A^T x A^T x A^T) v (interpolator_t%map to the coarse space) */
void orc_tnsr3d_t(double *u, int nu, const double *v, int nv, const double *A, int nelv);
void orc_opgrad(double *ux, double *uy, double *uz, const double *u, int lx, int nelv,
                const double *D, const double *w3, double *const G[9]);
void orc_cdtp(double *dtx, const double *x, const double *dr, const double *ds, const double *dt,
              int lx, int nelv, const double *D, const double *w3);
void orc_conv1(double *du, const double *u, const double *vx, const double *vy, const double *vz,
               int lx, int nelv, const double *D, double *const G[9], const double *jacinv);

/* ---- the hot path --------------------------------------------------------------------------- */
/* adv_adjoint_no_dealias.f90:119-255 (intended = device-branch semantics :162-201,269-303).
 * f is IN/OUT (accumulated).  bug_compat!=0 reproduces CPU-branch defects D1 (index shift,
 * :213-216) and D2 (sign, :344) for forensic comparison only. */
void orc_adjoint_advection_no_dealias(double *fx, double *fy, double *fz,
                                      const double *vx, const double *vy, const double *vz,
                                      const double *vxb, const double *vyb, const double *vzb,
                                      int lx, int nelv, const double *D, const double *w,
                                      double *const G[9], int bug_compat);
/* adv_adjoint_no_dealias.f90:365-427 (linearised operator; used for the adjoint identity) */
void orc_linear_advection_no_dealias(double *fx, double *fy, double *fz,
                                     const double *vx, const double *vy, const double *vz,
                                     const double *vxb, const double *vyb, const double *vzb,
                                     int lx, int nelv, const double *D, const double *w,
                                     double *const G[9], const double *jac);
/* adv_adjoint_dealias.f90:137-161 + 235-462 (CPU branch with per-element geometry).
 * lxd = fine GL order; GLL-space G is interpolated to GL as init_dealias does. */
void orc_adjoint_advection_dealias(double *fx, double *fy, double *fz,
                                   const double *vx, const double *vy, const double *vz,
                                   const double *vxb, const double *vyb, const double *vzb,
                                   int lx, int lxd, int nelv, double *const G[9]);
/* adv_adjoint_dealias.f90:479-668 (linearised operator, dealiased) */
void orc_linear_advection_dealias(double *fx, double *fy, double *fz,
                                  const double *vx, const double *vy, const double *vz,
                                  const double *vxb, const double *vyb, const double *vzb,
                                  int lx, int lxd, int nelv, double *const G[9]);

/* RAMP_mapping.f90:182-196 (convex down) / :227-241 (convex up) */
void orc_ramp(double *chi, const double *rho, int64_t n, double f_min, double f_max, double q,
              int convex_up);
/* RAMP_mapping.f90:203-222 / :248-267 chain rule */
void orc_ramp_backward(double *dF_drho, const double *dF_dchi, const double *rho, int64_t n,
                       double f_min, double f_max, double q, int convex_up);
/* simple_brinkman_source_term.f90:139-153: f_i -= chi*u_i */
void orc_brinkman(double *fx, double *fy, double *fz, const double *u, const double *v,
                  const double *w, const double *chi, int64_t n);
/* adjoint_lube_source_term.f90:173-206: f_i += (K*chi [masked]) * u_i; mask = 1-based indices
 * kept (mask_ops.f90:55-82), everything else zeroed; mask==NULL -> no mask */
void orc_lube(double *fx, double *fy, double *fz, const double *u, const double *v,
              const double *w, const double *chi, double K, const int *mask, int mask_size,
              int64_t n);
/* adjoint_pnpn.f90:671-676: f_i *= B */
void orc_opcolv(double *fx, double *fy, double *fz, const double *B, int64_t n);
/* minimum_dissipation_objective_function.f90:260-301 */
void orc_sensitivity(double *S, const double *u, const double *v, const double *w,
                     const double *ua, const double *va, const double *wa, double K_obj,
                     int if_lube, int64_t n);

/* ---- minimum-dissipation objective chain -------------------------------------------
 * Neko operators restated from their published CPU back-end (operators.f90 / opr_cpu_*; not vendored):
 * dudxyz: du = jacinv * (dr*ur + ds*us + dt*ut)  (strong derivative in one physical direction) */
void orc_dudxyz(double *du, const double *u, const double *dr, const double *ds, const double *dt,
                const double *jacinv, int lx, int nelv, const double *D);
/* curl(w, u): strong curl, then mass-weighted averaging over coincident nodes:
 *   w1 = du3/dy - du2/dz, w2 = du1/dz - du3/dx, w3 = du2/dx - du1/dy; w *= B; gs_op(w, ADD); w *= Binv
 * (Binv = 1/gs(B), coef_t).  class_id/nclass: orc_gs_classes of the mesh. */
void orc_curl(double *w1, double *w2, double *w3, const double *u1, const double *u2, const double *u3,
              int lx, int nelv, const double *D, double *const G[9], const double *jacinv,
              const double *B, const double *Binv, const int64_t *class_id, int64_t nclass);
/* mask_ops.f90:55-82 mask_exterior_const: keep fld on the 1-based mask indices, `c` elsewhere */
void orc_mask_exterior_const(double *fld, const int *mask, int mask_size, double c, int64_t n);
/* math_ext.f90:99-116 glsc2_mask (local part): sum_{i in mask} a_i*b_i; mask == NULL -> glsc2 */
double orc_glsc2_mask(const double *a, const double *b, const int *mask, int mask_size, int64_t n);
/* adjoint_minimum_dissipation_source_term.f90:175-249: f += obj_scale * curl(curl(u)) [masked] */
void orc_curlcurl_forcing(double *fu, double *fv, double *fw, const double *u, const double *v,
                          const double *w, int lx, int nelv, const double *D, double *const G[9],
                          const double *jacinv, const double *B, const double *Binv,
                          const int64_t *class_id, int64_t nclass, const int *mask, int mask_size,
                          double obj_scale);
/* minimum_dissipation_objective_function.f90:186-254: out[0] = dissipation = sum |grad u_c|^2 * B,
 * out[1] = lube_value = sum (u+v+w)*chi * B -- what :230-232 literally computes (col3/addcol3 with the
 * velocity components, not their squares; restated as written) -- 0 if chi == NULL; both over the mask if given;
 * returns the objective (dissipation + 0.5*K*lube_value)*obj_scale */
double orc_min_dissipation_objective(double out[2], const double *u, const double *v, const double *w,
                                     const double *chi, int lx, int nelv, const double *D,
                                     double *const G[9], const double *jacinv, const double *B,
                                     const int *mask, int mask_size, double K, double obj_scale);

/* ---- PDE filter: Neko ax_helm restated (not vendored) ---------------------------------
 * w = D^T (h1 G D u) + h2 B u per element, G_ij = sum_k (dr_i/dx_k)(dr_j/dx_k) * jacinv * w3 from the cofactors.
 * The filter itself, (r^2 K + M) x = gs(B x_in), is solved in tests/ by assembling this operator. */
void orc_ax_helm(double *w, const double *u, int lx, int nelv, const double *D, const double *wq,
                 double *const G[9], const double *jacinv, const double *B, double h1, double h2);

/* ---- explicit time scheme around the RHS (adjoint_pnpn.f90:665-666,688-696) -------------------
 * The three rhs_maker types live in Neko (src/fluid/rhs_maker*.f90, not vendored); restated from
 * Neko's published CPU back-end (rhs_maker_cpu.f90), argument order of the reference's call sites.
 * sumab%compute_fluid: u_e = ab(1)*u + ab(2)*ulag(1) [+ ab(3)*ulag(2) if nab == 3] */
void orc_sumab(double *ue, double *ve, double *we, const double *u, const double *v, const double *w,
               const double *ulag1, const double *vlag1, const double *wlag1,
               const double *ulag2, const double *vlag2, const double *wlag2,
               const double ab[3], int nab, int64_t n);
/* makeabf%compute_fluid: ta = ext(2)*f_lag + ext(3)*f_laglag; f_laglag = f_lag; f_lag = f;
 * f = (ext(1)*f + ta)*rho   (all three components; lag arrays updated in place) */
void orc_makeabf(double *abx1, double *aby1, double *abz1, double *abx2, double *aby2, double *abz2,
                 double *fx, double *fy, double *fz, double rho, const double ext[3], int64_t n);
/* makebdf%compute_fluid: tb = u*B*bd(2) + sum_{ilag=2..nbd} ulag(ilag-1)*B*bd(ilag+1);
 * f = f + tb*(rho/dt)   (bd has nbd+1 entries; nbd <= 3) */
void orc_makebdf(const double *ulag1, const double *vlag1, const double *wlag1,
                 const double *ulag2, const double *vlag2, const double *wlag2,
                 double *fx, double *fy, double *fz, const double *u, const double *v, const double *w,
                 const double *B, double rho, double dt, const double bd[4], int nbd, int64_t n);

/* The whole explicit-RHS slice adjoint_pnpn.f90:661-682:
 *   f = 0; brinkman(u_adj, chi); [lube(u_b, chi)]; [f += f_static]; f *= B; adv%compute_adjoint.
 * chi = RAMP(rho) if chi_in == NULL.  sens may be NULL.  lxd == 0 -> no dealias. */
typedef struct {
  double f_min, f_max, q;
  int convex_up;
  int if_lube;
  double K_lube;  /* K*obj_scale handed to the lube term (min_diss_objective :158-166) */
  double K_sens;  /* K*obj_scale in the sensitivity (:295-296) */
  int lxd;
} orc_params;
void orc_adjoint_rhs(double *fx, double *fy, double *fz, double *sens, double *chi_out,
                     const double *vx, const double *vy, const double *vz,
                     const double *vxb, const double *vyb, const double *vzb,
                     const double *rho, const double *chi_in,
                     const double *fsx, const double *fsy, const double *fsz,
                     const int *mask, int mask_size,
                     int lx, int nelv, const double *D, const double *w,
                     double *const G[9], const double *B, const orc_params *p);

/* ---- gather-scatter (Neko gs_t%op(., GS_OP_ADD); adjoint_pnpn.f90:725,755-757) -------------- */
/* class_id[n]: canonical relabelling -- classes numbered by first appearance (ascending dof).
 * returns number of classes. */
int64_t orc_gs_classes(int64_t *class_id, const int64_t *key, int64_t n);
/* in-place direct-stiffness sum; members summed in ascending dof order */
void orc_gs_add(double *f, const int64_t *class_id, int64_t nclass, int64_t n);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
