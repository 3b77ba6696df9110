/*
 * neko_top_b200.h -- C ABI of the B200-native adjoint-RHS path for Neko-TOP.
 *
 * Drop-in boundary for SURVEY.md section 8 rows (a1)-(a12): every entry point is `extern "C"`,
 * takes plain pointers and sizes (no C++ / torch types), and follows the calling convention of the
 * reference's own CUDA shims so that a Fortran `bind(c)` interface block binds it directly:
 *   - device pointers are passed BY VALUE as `void*`   (Fortran: `type(c_ptr), value`)
 *   - scalars are passed BY REFERENCE                  (Fortran: `integer(c_int) :: n`, `real(c_rp) :: c`)
 *   - masks are 1-based linear indices                 (math_ext_kernel.h:49 `a[mask[i]-1]`)
 * exactly as /root/reference/sources/neko_ext/math/bcknd/device_math_ext.f90:43-103 binds
 * /root/reference/sources/neko_ext/math/bcknd/device/cuda/math_ext.cu:49-108.
 * Fields are fp64 (Neko `rp` = `dp`), Fortran column-major x(lx,lx,lx,nelv); device field pointers must be
 * 16-byte aligned (any cudaMalloc / Neko device_alloc pointer is).  For odd lx the element kernels fetch
 * k-planes with 16-byte aligned TMA windows and may READ (never write) up to 8 bytes past the last double
 * of an input field, inside the same 16-byte granule -- always inside the allocation.
 *
 * Error convention (reference: neko_error / CUDA_CHECK abort the job, math_ext.cu:56): every call
 * returns 0 on success; on failure it prints the reason to stderr and, unless
 * b200_set_abort_on_error(0) was called, aborts -- there is no CPU fallback anywhere.
 *
 * All citations are relative to /root/reference/sources.
 */
#ifndef NEKO_TOP_B200_H
#define NEKO_TOP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_ARG 1
#define B200_ERR_CUDA 2
#define B200_ERR_NCCL 3
#define B200_ERR_STATE 4

/* library-wide */
int b200_version(void);
void b200_set_abort_on_error(const int* flag);
const char* b200_last_error(void);
/* number of kernels launched by this library since load (bench.py "gpu_launches") */
int64_t b200_launch_count(void);

/* ---- handle: one per (space, mesh partition) == one per Neko `coef_t` -----------------------
 * replaces the state held by adv_lin_no_dealias_t (adjoint/adv_adjoint_no_dealias.f90:57-77)
 * and adv_lin_dealias_t (adjoint/adv_adjoint_dealias.f90:56-131). */
/* *device: CUDA device index, or -1 (or NULL) = the calling thread's current device (what a Neko rank selected
 * in device_init; use this from Fortran).  Every entry point makes the handle's device current. */
int b200_adjrhs_create(void** handle, const int* lx, const int* nelv, const int* device);
int b200_adjrhs_free(void** handle);
/* Neko enqueues everything on glb_cmd_queue (math_ext.cu:54); pass it here (NULL = default). */
int b200_adjrhs_set_stream(void* handle, void* stream);
/* space_t: dx = Xh%dx (lx*lx, column-major D(i,j)), wx = Xh%wx (lx); HOST pointers. */
int b200_adjrhs_set_space(void* handle, const double* dx, const double* wx);
/* coef_t device mirrors: coef%drdx_d ... coef%dtdz_d (cofactors, J-scaled) and coef%B_d.
 * Borrowed; must stay valid while the handle is used.  The call also builds (on the handle's stream,
 * synchronously) a private per-plane interleaved image of the ten arrays for the fused kernel -- one
 * TMA bulk copy per k-plane instead of ten -- which costs 10*n*8 bytes of device memory; geometry is
 * constant during a Neko-TOP run (coef_t is built once, adjoint_scheme.f90:341-343), call it again if
 * the arrays ever change.  Call b200_adjrhs_set_stream first. */
int b200_adjrhs_set_geometry(void* handle,
                             const void* drdx, const void* dsdx, const void* dtdx,
                             const void* drdy, const void* dsdy, const void* dtdy,
                             const void* drdz, const void* dsdz, const void* dtdz,
                             const void* B);
/* RAMP constants (mapping_functions/RAMP_mapping.f90:107-110), lube K*obj_scale and the
 * sensitivity K*obj_scale (objectives/minimum_dissipation_objective_function.f90:131-133). */
int b200_adjrhs_set_params(void* handle, const double* f_min, const double* f_max, const double* q,
                           const int* convex_up, const int* if_lube, const double* K_lube,
                           const double* K_sens);
/* optional point-zone mask of the lube term (source_terms/adjoint_lube_source_term.f90:196-198,
 * neko_ext/mask_ops.f90:55-82): DEVICE array of 1-based indices; mask_size = 0 disables. */
int b200_adjrhs_set_lube_mask(void* handle, const void* mask_d, const int* mask_size);

/* ---- the fused path: adjoint/adjoint_pnpn.f90:669-682 in one pass per element ---------------
 *   chi = RAMP(rho)            (rho != NULL; else chi taken from chi_in; both NULL: no source terms)
 *   f_i = B*(-chi*v_i [+K_lube*chi*vb_i] [+fs_i]) - adjoint advection(v, vb)
 *   sens = -(vb.v) + K_sens*(vb.vb)        (sens != NULL)
 * f is WRITE-ONLY here.  All pointers are device pointers; optional ones may be NULL. */
int b200_adjrhs_compute(void* handle,
                        const void* vx, const void* vy, const void* vz,
                        const void* vxb, const void* vyb, const void* vzb,
                        const void* rho, const void* chi_in,
                        const void* fsx, const void* fsy, const void* fsz,
                        void* fx, void* fy, void* fz,
                        void* sens, void* chi_out);
/* fused path followed by gs_op(f_i, GS_OP_ADD) on the three components
 * (adjoint_pnpn.f90:755-757); needs b200_gs_init.  This is one "step" of bench.py. */
int b200_adjrhs_step(void* handle,
                     const void* vx, const void* vy, const void* vz,
                     const void* vxb, const void* vyb, const void* vzb,
                     const void* rho, const void* chi_in,
                     const void* fsx, const void* fsy, const void* fsz,
                     void* fx, void* fy, void* fz,
                     void* sens, void* chi_out);
/* same step with HOST buffers (each n doubles): H2D of v, vb, rho; D2H of f, sens; staged through
 * pinned memory in element chunks so copies overlap the kernels.  Geometry stays device-resident
 * (it is set once, like coef_t).  This is bench.py's "e2e". */
int b200_adjrhs_step_host(void* handle,
                          const double* vx, const double* vy, const double* vz,
                          const double* vxb, const double* vyb, const double* vzb,
                          const double* rho,
                          double* fx, double* fy, double* fz, double* sens);

/* ---- un-fused drop-ins, one per reference plug-in method -----------------------------------
 * advection_adjoint_t%compute_adjoint (adjoint/advection_adjoint.f90:67-81; implementation
 * adjoint/adv_adjoint_no_dealias.f90:119-255): f is IN/OUT, accumulated. */
int b200_adv_adjoint_compute(void* handle,
                             const void* vx, const void* vy, const void* vz,
                             const void* vxb, const void* vyb, const void* vzb,
                             void* fx, void* fy, void* fz);
/* advection_adjoint_t%compute_linear on the GLL grid (adjoint/adv_adjoint_no_dealias.f90:365-427):
 * f_i -= B*jacinv*(U_b.grad u'_i + u'.grad U_b,i); f IN/OUT.  jacinv (coef%jacinv_d) is accepted for
 * signature compatibility and may be NULL: B*jacinv == w3, which the kernel uses directly. */
int b200_adv_linear_compute(void* handle,
                            const void* vx, const void* vy, const void* vz,
                            const void* vxb, const void* vyb, const void* vzb,
                            const void* jacinv,
                            void* fx, void* fy, void* fz);
/* adv_lin_dealias_t%init (adjoint/adv_adjoint_dealias.f90:137-161).  HOST arrays of the fine
 * Gauss-Legendre space: *lxd = Xh_GL%lx, interp = GLL_to_GL interpolation matrix (lxd x lx, column-major
 * J(a,l)), dxd = Xh_GL%dx (lxd x lxd, column-major), wd = Xh_GL%wx (lxd).  Interpolates the nine
 * geometric factors set by b200_adjrhs_set_geometry to the fine grid (coef_GL, :153-161) into
 * library-owned memory (9*nelv*lxd^3*8 bytes).  Only lxd = 3*lx/2 -- the factory's default,
 * adjoint/advection_adjoint_fctry.f90:70,89 -- is instantiated; anything else is an error. */
int b200_adv_dealias_init(void* handle, const int* lxd, const double* interp, const double* dxd,
                          const double* wd);
/* adv_lin_dealias_t%compute_adjoint (adjoint/adv_adjoint_dealias.f90:235-462); f IN/OUT.
 * lx = 8 / lxd = 12 runs on the FP64 tensor cores (csrc/advop_mma_kernel.cuh; environment B200_ADVOP_MMA=0 selects
 * the column-per-thread kernel the other orders use); same interface, same 1e-12 parity. */
int b200_adv_adjoint_dealias_compute(void* handle,
                                     const void* vx, const void* vy, const void* vz,
                                     const void* vxb, const void* vyb, const void* vzb,
                                     void* fx, void* fy, void* fz);
/* adv_lin_dealias_t%compute_linear (adjoint/adv_adjoint_dealias.f90:479-668); f IN/OUT. */
int b200_adv_linear_dealias_compute(void* handle,
                                    const void* vx, const void* vy, const void* vz,
                                    const void* vxb, const void* vyb, const void* vzb,
                                    void* fx, void* fy, void* fz);
/* *flag != 0: b200_adjrhs_compute / _step / _step_host evaluate the adjoint advection with the
 * dealiased operator (case.numerics.dealias = true) instead of the GLL-grid one; needs
 * b200_adv_dealias_init.  b200_adv_adjoint_compute always stays on the GLL grid. */
int b200_adjrhs_set_dealias(void* handle, const int* flag);
/* simple_brinkman_source_term_t%compute_ (source_terms/simple_brinkman_source_term.f90:139-153):
 * f_i -= chi*u_i */
int b200_brinkman_compute(void* fu, void* fv, void* fw, const void* u, const void* v, const void* w,
                          const void* chi, const int* n, void* stream);
/* adjoint_lube_source_term_t%compute_ (source_terms/adjoint_lube_source_term.f90:173-206):
 * f_i += K*chi*u_i on the mask (mask_size = 0: everywhere) */
int b200_lube_compute(void* fu, void* fv, void* fw, const void* u, const void* v, const void* w,
                      const void* chi, const double* K, const void* mask_d, const int* mask_size,
                      const int* n, void* stream);
/* opcolv (adjoint/adjoint_pnpn.f90:672-676): f_i *= B */
int b200_opcolv(void* fx, void* fy, void* fz, const void* B, const int* n, void* stream);
/* RAMP_mapping_t%apply_forward / apply_backward (mapping_functions/RAMP_mapping.f90:137-267) */
int b200_ramp_forward(void* chi, const void* rho, const int* n, const double* f_min,
                      const double* f_max, const double* q, const int* convex_up, void* stream);
int b200_ramp_backward(void* dF_drho, const void* dF_dchi, const void* rho, const int* n,
                       const double* f_min, const double* f_max, const double* q,
                       const int* convex_up, void* stream);
/* minimum_dissipation_objective_function_t%compute_sensitivity
 * (objectives/minimum_dissipation_objective_function.f90:260-301) */
int b200_sensitivity(void* sens, const void* u, const void* v, const void* w, const void* ua,
                     const void* va, const void* wa, const double* K_obj, const int* if_lube,
                     const int* n, void* stream);
/* steady_simcomp_t%compute_ (simulation_components/steady_simcomp.f90:158-188), one field per call:
 * *result = sum_i (x_old - x)^2 (the local part of field_sub2 + field_glsc2) and x_old <- x
 * (field_copy), in one pass.  Synchronises the stream (it returns a host scalar). */
int b200_steady_field_update(double* result, const void* x, void* x_old, const int* n, void* stream);

/* ---- minimum-dissipation objective chain ------------------------------------------
 * Neko `curl(w1,w2,w3, u1,u2,u3, work1, work2, coef)` as called at
 * source_terms/adjoint_minimum_dissipation_source_term.f90:231-232: strong derivatives (dudxyz), then
 * w *= B, gs_op(w, GS_OP_ADD), w *= Binv.  jacinv = coef%jacinv_d, Binv = coef%Binv_d; needs b200_gs_init. */
int b200_curl(void* handle, void* w1, void* w2, void* w3, const void* u1, const void* u2, const void* u3,
              const void* jacinv, const void* Binv);
/* adjoint_minimum_dissipation_source_term_t%compute_ (:175-249): f += obj_scale * curl(curl(u)), restricted to
 * the 1-based point-zone mask when *mask_size > 0 (mask_exterior_const(., mask, 0) + field_add2s2).
 * The six scratch fields of :202-209 are owned by the handle. */
int b200_curlcurl_forcing(void* handle, void* fu, void* fv, void* fw, const void* u, const void* v,
                          const void* w, const void* jacinv, const void* Binv, const void* mask_d,
                          const int* mask_size, const double* obj_scale);
/* minimum_dissipation_objective_function_t%compute (objectives/minimum_dissipation_objective_function.f90:
 * 186-254), rank-local part (the reference's glsc2 adds an MPI_Allreduce): out3[1] = dissipation =
 * sum_c |grad u_c|^2 . B, out3[2] = lube_value = ((u+v+w)*chi) . B as written at :230-232 (0 if chi == NULL),
 * both over the mask if *mask_size > 0; out3[0] = (dissipation + 0.5*K*lube_value)*obj_scale.
 * Synchronises the handle's stream (host scalars); sums are deterministic. */
int b200_min_dissipation_objective(void* handle, const void* u, const void* v, const void* w, const void* chi,
                                   const void* jacinv, const void* mask_d, const int* mask_size,
                                   const double* K, const double* obj_scale, double* out3);
/* mask_exterior_const (neko_ext/mask_ops.f90:55-82) on the device -- the reference errors out there
 * ("GPU not supported for masks yet", :70): fld keeps its values on the 1-based mask, *c elsewhere;
 * work = scratch field of *n doubles (Neko's scratch registry). */
int b200_mask_exterior_const(void* fld, void* work, const void* mask_d, const int* mask_size, const double* c,
                             const int* n, void* stream);

/* ---- PDE (Helmholtz) filter ------------------------------------------------------------
 * PDE_filter_t%apply / %apply_backward (mapping_functions/PDE_filter_mapping.f90:212-363): solve
 *   (r^2 K + M) x = gs(B * x_in)        (coef%h1 = r^2, coef%h2 = 1, ifh2; no Dirichlet conditions)
 * with Neko's preconditioned conjugate gradients (ax_helm, gs_op on every operator application, inner products
 * weighted with coef%mult, stop when sqrt(r.mult.r)*norm_fac < abs_tol).  The reference selects "gmres" + "ident"
 * (:133-137); the operator is symmetric positive definite, so CG converges to the same solution within the same
 * tolerance -- only ksp_results%iter differs (documented deviation).  Both directions of the filter are
 * this call (the backward pass filters the sensitivity with the same operator).
 * jacinv = coef%jacinv_d, mult = coef%mult_d; *precond: 0 = ident (the reference's choice), otherwise jacobi
 * (diagonal without the cross terms of deformed elements); norm_fac = 1/sqrt(volume) as in Neko, NULL = 1;
 * *x0_is_input != 0: start from x = x_in ("copy the unfiltered design as an initial guess", :246-248), 0 or NULL:
 * from x = 0 (what Neko's Krylov solvers do on entry).  The whole iteration runs on the device: inner products
 * are deterministic two-stage reductions into a device scalar (ncclAllReduce over the ranks in a multi-GPU run),
 * alpha / beta never visit the host, and the host looks at the residual once every 8 iterations (kernels past
 * convergence are no-ops, so the result is that of stopping at the exact iteration).  Needs b200_gs_init. */
int b200_pde_filter_apply(void* handle, void* x_out, const void* x_in, const void* jacinv, const void* mult,
                          const double* radius, const double* abs_tol, const int* max_iter,
                          const int* precond, const double* norm_fac, const int* x0_is_input, int* iters,
                          double* res_start, double* res_final);

/* ---- explicit time scheme around the RHS (adjoint/adjoint_pnpn.f90:665-666,688-696) --
 * Neko's rhs_maker types, argument order of the reference's call sites; all fields are device pointers of
 * *n doubles; coefficient arrays are HOST pointers.
 * sumab%compute_fluid(u_e,v_e,w_e, u,v,w, ulag,vlag,wlag, ext_bdf%advection_coeffs, nadv):
 *   u_e = ab(1)*u + ab(2)*ulag(1) [+ ab(3)*ulag(2) if *nab == 3]; ulag1 = ulag%lf(1)%x_d, ulag2 = ulag%lf(2)%x_d */
int b200_sumab(void* ue, void* ve, void* we, const void* u, const void* v, const void* w,
               const void* ulag1, const void* vlag1, const void* wlag1,
               const void* ulag2, const void* vlag2, const void* wlag2,
               const double* ab, const int* nab, const int* n, void* stream);
/* makeabf%compute_fluid(abx1,aby1,abz1, abx2,aby2,abz2, f_x,f_y,f_z, rho, advection_coeffs, n):
 *   ta = ext(2)*ab1 + ext(3)*ab2; ab2 = ab1; ab1 = f; f = (ext(1)*f + ta)*rho */
int b200_makeabf(void* abx1, void* aby1, void* abz1, void* abx2, void* aby2, void* abz2,
                 void* fx, void* fy, void* fz, const double* rho, const double* ext, const int* n,
                 void* stream);
/* makebdf%compute_fluid(ulag,vlag,wlag, f_x,f_y,f_z, u,v,w, B, rho, dt, diffusion_coeffs, ndiff, n):
 *   tb = u*B*bd(2) + sum_{ilag=2..nbd} ulag(ilag-1)*B*bd(ilag+1); f = f + tb*(rho/dt); bd has *nbd+1 entries */
int b200_makebdf(const void* ulag1, const void* vlag1, const void* wlag1,
                 const void* ulag2, const void* vlag2, const void* wlag2,
                 void* fx, void* fy, void* fz, const void* u, const void* v, const void* w,
                 const void* B, const double* rho, const double* dt, const double* bd, const int* nbd,
                 const int* n, void* stream);
/* both in ONE pass over f (the two calls above back to back, adjoint_pnpn.f90:688-696) */
int b200_makeabf_bdf(void* abx1, void* aby1, void* abz1, void* abx2, void* aby2, void* abz2,
                     const void* ulag1, const void* vlag1, const void* wlag1,
                     const void* ulag2, const void* vlag2, const void* wlag2,
                     void* fx, void* fy, void* fz, const void* u, const void* v, const void* w,
                     const void* B, const double* rho, const double* dt, const double* ext,
                     const double* bd, const int* nbd, const int* n, void* stream);

/* ---- gather-scatter: gs_t%op(., GS_OP_ADD) (adjoint/adjoint_pnpn.f90:725,755-757) ------------
 * key: global node id of every local dof (n = nelv*lx^3 int64), device pointer if *on_device.
 * Two dofs are summed iff their keys are equal (SURVEY.md 8c "node equivalence classes"). */
int b200_gs_init(void* handle, const int64_t* key, const int* on_device);
/* parity hook: class id per dof, classes numbered by ascending smallest member dof (the same
 * canonical relabelling the oracle uses); HOST output, n int64. Returns #classes in *nclass. */
int b200_gs_get_classes(void* handle, int64_t* class_id, int64_t* nclass);
int b200_gs_op(void* handle, void* f);                       /* one field   */
int b200_gs_op3(void* handle, void* fx, void* fy, void* fz); /* three fields, one pass */

/* ---- multi-GPU: one process per GPU, shared-node exchange as NCCL send/recv over NVLink -------
 * id: 128-byte ncclUniqueId made by rank 0 (b200_comm_unique_id) and distributed by the host
 * (MPI_Bcast in Neko, torch.distributed in bench.py). */
int b200_comm_unique_id(char* id128);
int b200_comm_init(void* handle, const char* id128, const int* rank, const int* nranks);
/* shared nodes, as Neko's gs_t stores them (shared dofs + per-neighbour lists):
 *   nshared          number of local node classes that also live on other ranks
 *   shared_dof[s]    0-based local dof index of any member of class s
 *   nneigh, neigh_rank[j], neigh_off[j..j+1], neigh_idx[...]: for neighbour j the indices s (into
 *   shared_dof) of the nodes shared with it, in an order both sides agree on (ascending key). */
int b200_gs_init_shared(void* handle, const int* nshared, const int* shared_dof, const int* nneigh,
                        const int* neigh_rank, const int* neigh_off, const int* neigh_idx);
/* Shared-node discovery behind the C ABI: what Neko's gs_t%init does from the dofmap
 * (adjoint/adjoint_scheme.f90:339-343 `gs_Xh%init(dm_Xh)`).  Collective over the communicator of b200_comm_init;
 * call after b200_gs_init with the SAME keys.  `cand` (n bytes, same memory space as `key`, may be NULL) flags
 * the dofs that can live on another rank -- from Fortran: Neko's dm_Xh%shared_dof; NULL = every dof on the
 * surface of its element (fine for small meshes, wasteful for large ones).  On the device: candidate keys are
 * sorted and made unique (CUB), all-gathered with ncclAllGather, and every rank intersects its list with every
 * other rank's; the per-neighbour message layout is the intersection in ascending key order, so both sides agree
 * without a handshake.  Ends by calling b200_gs_init_shared and b200_adjrhs_set_boundary_elements with the lists
 * it found.  *nshared / *nneigh (optional) return the number of shared nodes / neighbour ranks. */
int b200_gs_init_shared_from_keys(void* handle, const int64_t* key, const int* on_device,
                                  const unsigned char* cand, int* nshared, int* nneigh);
/* elements that own a shared node (computed first so the exchange overlaps the interior ones) */
int b200_adjrhs_set_boundary_elements(void* handle, const int* nbnd, const int* bnd_elem);

/* Processing order of the elements in the fused step (a permutation of 0..nelv-1, HOST; *nelem = 0
 * restores 0..nelv-1).  Results do not depend on it.  The staged summation and the contiguous-run element
 * kernel need the mesh order, so any order set here switches them off (and the element-list variant of the
 * kernel is ~10 % slower): a diagnostic / experimentation hook. */
int b200_adjrhs_set_element_order(void* handle, const int* nelem, const int* order);
/* Class-list pass of b200_adjrhs_step when the staged summation (b200_adjrhs_set_xstage) is not in use:
 *   *flag = 0 (default)  CSR class lists;
 *   *flag != 0           class lists packed by size, sorted by the element that completes them (the lists the
 *                        pipelined b200_adjrhs_step_host walks; environment B200_GS_MODE=1).  Same bits either way.
 * (Round 1's experimental summation inside the element kernel through completion counters is gone: the staged
 * summation removes the same DRAM traffic without spin-waits.) */
int b200_adjrhs_set_gs_fused(void* handle, const int* flag);
/* *fused = 0 (kept for ABI stability); *classes_in_kernel = classes in the packed lists when they are in use */
int b200_adjrhs_gs_info(void* handle, int* fused, int64_t* classes_in_kernel, int64_t* classes_total);
/* Staged direct-stiffness summation of b200_adjrhs_step at lx = 8 (environment B200_XSTAGE):
 *   *flag = 0  plain element kernel + gather-scatter pass over all class lists;
 *   *flag = 1  every element slot walks a contiguous run of elements; where consecutive elements e-1, e are glued
 *              i = lx-1 -> i = 0 node by node the kernel sums the 2-member classes of that face itself (the
 *              partner value is re-read from L2) and the pass runs over the remaining classes -- it then touches
 *              44 % instead of 100 % of the 32-byte sectors of f; bit-identical to flag 0 (a + b commutes);
 *   *flag = 2  (default) additionally every class that is a PRODUCT of face pairings (face-interior nodes, 2x2
 *              edges, 2x2x2 vertices of a locally structured mesh) is summed direction by direction: x in the
 *              kernel, y and z by two face passes over contiguous rows / planes; only irregular and
 *              partition-boundary classes stay in the class-list pass.  All copies of a node still end with
 *              identical bits; 4- and 8-member sums are associated pairwise, i.e. they differ from flag 0 by
 *              rounding (<= 1e-15 relative).
 * Everything is verified against the classes of b200_gs_init at set-up; nothing is assumed about the mesh
 * (gs_kernels.cuh "staged").  Used when the elements are in mesh order (no b200_adjrhs_set_element_order), gs
 * mode 0, no point-zone mask.  Replaces gs_Xh%op(f, GS_OP_ADD) after the RHS (adjoint_pnpn.f90:755-757). */
int b200_adjrhs_set_xstage(void* handle, const int* flag);
/* *active = level used by the last/next b200_adjrhs_step (0, 1, 2); *classes_staged = classes summed by the
 * kernel / face passes; *classes_left of *classes_total stay in the class-list pass */
int b200_adjrhs_xstage_info(void* handle, int* active, int64_t* classes_staged, int64_t* classes_left,
                            int64_t* classes_total);

/* ---- diagnostics ----------------------------------------------------------------------------*/
/* average device time (ms) of the last fused element-kernel launches measured with CUDA events
 * on the handle's stream when profiling is enabled; used by bench.py's roofline block */
int b200_adjrhs_enable_timing(void* handle, const int* flag);
int b200_adjrhs_get_timing(void* handle, double* elem_kernel_ms, double* gs_ms, int64_t* launches);
/* with environment B200_PHASE_TIMING=1: device time (ms) of the phases of the LAST multi-GPU step --
 * boundary elements, shared-class gs, pack + exchange issue, interior elements, local gs, unpack;
 * in: *nphase = capacity of ms[], out: number of phases */
int b200_adjrhs_get_phase_timing(void* handle, double* ms, int* nphase);

#ifdef __cplusplus
}
#endif
#endif
