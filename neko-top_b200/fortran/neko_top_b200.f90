!> iso_c_binding interface of libneko_top_b200.so (include/neko_top_b200.h).
!!
!! Follows the reference's own CUDA-shim convention
!! (sources/neko_ext/math/bcknd/device_math_ext.f90:43-103): device pointers are
!! `type(c_ptr), value`, scalars are passed by reference.  Every function returns
!! an integer status (0 = ok); by default the library aborts on error like
!! `neko_error` / `CUDA_CHECK` do (math_ext.cu:56).  The shim still passes every
!! status through `b200_check`, which ends the job with `neko_error` and the
!! library's message if abort-on-error was switched off.
!!
!! NOTE: this file could not be compiled in the build image (no Fortran compiler,
!! SURVEY.md 0.3); it is checked textually against the C header by
!! tests/test_abi.py::test_fortran_shim_binds_every_symbol.
module neko_top_b200
  use, intrinsic :: iso_c_binding
  implicit none
  public

  interface
     ! ---- library-wide -----------------------------------------------------
     integer(c_int) function b200_version() bind(c, name='b200_version')
       use, intrinsic :: iso_c_binding
     end function b200_version

     subroutine b200_set_abort_on_error(flag) &
          bind(c, name='b200_set_abort_on_error')
       use, intrinsic :: iso_c_binding
       integer(c_int) :: flag
     end subroutine b200_set_abort_on_error

     type(c_ptr) function b200_last_error() bind(c, name='b200_last_error')
       use, intrinsic :: iso_c_binding
     end function b200_last_error

     integer(c_int64_t) function b200_launch_count() &
          bind(c, name='b200_launch_count')
       use, intrinsic :: iso_c_binding
     end function b200_launch_count

     ! ---- handle -----------------------------------------------------------
     integer(c_int) function b200_adjrhs_create(handle, lx, nelv, device) &
          bind(c, name='b200_adjrhs_create')
       use, intrinsic :: iso_c_binding
       type(c_ptr) :: handle
       integer(c_int) :: lx, nelv, device
     end function b200_adjrhs_create

     integer(c_int) function b200_adjrhs_free(handle) &
          bind(c, name='b200_adjrhs_free')
       use, intrinsic :: iso_c_binding
       type(c_ptr) :: handle
     end function b200_adjrhs_free

     integer(c_int) function b200_adjrhs_set_stream(handle, stream) &
          bind(c, name='b200_adjrhs_set_stream')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle, stream
     end function b200_adjrhs_set_stream

     integer(c_int) function b200_adjrhs_set_space(handle, dx, wx) &
          bind(c, name='b200_adjrhs_set_space')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       real(c_double) :: dx(*), wx(*)
     end function b200_adjrhs_set_space

     integer(c_int) function b200_adjrhs_set_geometry(handle, &
          drdx_d, dsdx_d, dtdx_d, drdy_d, dsdy_d, dtdy_d, &
          drdz_d, dsdz_d, dtdz_d, B_d) bind(c, name='b200_adjrhs_set_geometry')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: drdx_d, dsdx_d, dtdx_d, drdy_d, dsdy_d, dtdy_d
       type(c_ptr), value :: drdz_d, dsdz_d, dtdz_d, B_d
     end function b200_adjrhs_set_geometry

     integer(c_int) function b200_adjrhs_set_params(handle, f_min, f_max, q, &
          convex_up, if_lube, K_lube, K_sens) &
          bind(c, name='b200_adjrhs_set_params')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       real(c_double) :: f_min, f_max, q, K_lube, K_sens
       integer(c_int) :: convex_up, if_lube
     end function b200_adjrhs_set_params

     integer(c_int) function b200_adjrhs_set_lube_mask(handle, mask_d, &
          mask_size) bind(c, name='b200_adjrhs_set_lube_mask')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle, mask_d
       integer(c_int) :: mask_size
     end function b200_adjrhs_set_lube_mask

     ! ---- the fused path ---------------------------------------------------
     integer(c_int) function b200_adjrhs_compute(handle, vx_d, vy_d, vz_d, &
          vxb_d, vyb_d, vzb_d, rho_d, chi_d, fsx_d, fsy_d, fsz_d, &
          fx_d, fy_d, fz_d, sens_d, chi_out_d) &
          bind(c, name='b200_adjrhs_compute')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: vx_d, vy_d, vz_d, vxb_d, vyb_d, vzb_d
       type(c_ptr), value :: rho_d, chi_d, fsx_d, fsy_d, fsz_d
       type(c_ptr), value :: fx_d, fy_d, fz_d, sens_d, chi_out_d
     end function b200_adjrhs_compute

     integer(c_int) function b200_adjrhs_step(handle, vx_d, vy_d, vz_d, &
          vxb_d, vyb_d, vzb_d, rho_d, chi_d, fsx_d, fsy_d, fsz_d, &
          fx_d, fy_d, fz_d, sens_d, chi_out_d) &
          bind(c, name='b200_adjrhs_step')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: vx_d, vy_d, vz_d, vxb_d, vyb_d, vzb_d
       type(c_ptr), value :: rho_d, chi_d, fsx_d, fsy_d, fsz_d
       type(c_ptr), value :: fx_d, fy_d, fz_d, sens_d, chi_out_d
     end function b200_adjrhs_step

     integer(c_int) function b200_adjrhs_step_host(handle, vx, vy, vz, &
          vxb, vyb, vzb, rho, fx, fy, fz, sens) &
          bind(c, name='b200_adjrhs_step_host')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       real(c_double) :: vx(*), vy(*), vz(*), vxb(*), vyb(*), vzb(*), rho(*)
       real(c_double) :: fx(*), fy(*), fz(*), sens(*)
     end function b200_adjrhs_step_host

     ! ---- un-fused drop-ins ------------------------------------------------
     integer(c_int) function b200_adv_adjoint_compute(handle, vx_d, vy_d, &
          vz_d, vxb_d, vyb_d, vzb_d, fx_d, fy_d, fz_d) &
          bind(c, name='b200_adv_adjoint_compute')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: vx_d, vy_d, vz_d, vxb_d, vyb_d, vzb_d
       type(c_ptr), value :: fx_d, fy_d, fz_d
     end function b200_adv_adjoint_compute

     integer(c_int) function b200_adv_linear_compute(handle, vx_d, vy_d, &
          vz_d, vxb_d, vyb_d, vzb_d, jacinv_d, fx_d, fy_d, fz_d) &
          bind(c, name='b200_adv_linear_compute')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: vx_d, vy_d, vz_d, vxb_d, vyb_d, vzb_d, jacinv_d
       type(c_ptr), value :: fx_d, fy_d, fz_d
     end function b200_adv_linear_compute

     integer(c_int) function b200_adv_dealias_init(handle, lxd, interp, dxd, wd) &
          bind(c, name='b200_adv_dealias_init')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       integer(c_int) :: lxd
       real(c_double), dimension(*) :: interp, dxd, wd
     end function b200_adv_dealias_init

     integer(c_int) function b200_adv_adjoint_dealias_compute(handle, vx_d, &
          vy_d, vz_d, vxb_d, vyb_d, vzb_d, fx_d, fy_d, fz_d) &
          bind(c, name='b200_adv_adjoint_dealias_compute')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: vx_d, vy_d, vz_d, vxb_d, vyb_d, vzb_d
       type(c_ptr), value :: fx_d, fy_d, fz_d
     end function b200_adv_adjoint_dealias_compute

     integer(c_int) function b200_adv_linear_dealias_compute(handle, vx_d, &
          vy_d, vz_d, vxb_d, vyb_d, vzb_d, fx_d, fy_d, fz_d) &
          bind(c, name='b200_adv_linear_dealias_compute')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: vx_d, vy_d, vz_d, vxb_d, vyb_d, vzb_d
       type(c_ptr), value :: fx_d, fy_d, fz_d
     end function b200_adv_linear_dealias_compute

     integer(c_int) function b200_adjrhs_get_phase_timing(handle, ms, nphase) &
          bind(c, name='b200_adjrhs_get_phase_timing')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       real(c_double), dimension(*) :: ms
       integer(c_int) :: nphase
     end function b200_adjrhs_get_phase_timing

     integer(c_int) function b200_curl(handle, w1_d, w2_d, w3_d, u1_d, u2_d, &
          u3_d, jacinv_d, Binv_d) bind(c, name='b200_curl')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: w1_d, w2_d, w3_d, u1_d, u2_d, u3_d, jacinv_d, Binv_d
     end function b200_curl

     integer(c_int) function b200_curlcurl_forcing(handle, fu_d, fv_d, fw_d, &
          u_d, v_d, w_d, jacinv_d, Binv_d, mask_d, mask_size, obj_scale) &
          bind(c, name='b200_curlcurl_forcing')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: fu_d, fv_d, fw_d, u_d, v_d, w_d, jacinv_d, Binv_d
       type(c_ptr), value :: mask_d
       integer(c_int) :: mask_size
       real(c_double) :: obj_scale
     end function b200_curlcurl_forcing

     integer(c_int) function b200_min_dissipation_objective(handle, u_d, v_d, &
          w_d, chi_d, jacinv_d, mask_d, mask_size, K, obj_scale, out3) &
          bind(c, name='b200_min_dissipation_objective')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: u_d, v_d, w_d, chi_d, jacinv_d, mask_d
       integer(c_int) :: mask_size
       real(c_double) :: K, obj_scale
       real(c_double), dimension(3) :: out3
     end function b200_min_dissipation_objective

     integer(c_int) function b200_mask_exterior_const(fld_d, work_d, mask_d, &
          mask_size, c, n, stream) bind(c, name='b200_mask_exterior_const')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: fld_d, work_d, mask_d
       integer(c_int) :: mask_size, n
       real(c_double) :: c
       type(c_ptr), value :: stream
     end function b200_mask_exterior_const

     integer(c_int) function b200_pde_filter_apply(handle, x_out_d, x_in_d, &
          jacinv_d, mult_d, radius, abs_tol, max_iter, precond, norm_fac, &
          x0_is_input, iters, res_start, res_final) &
          bind(c, name='b200_pde_filter_apply')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: x_out_d, x_in_d, jacinv_d, mult_d
       real(c_double) :: radius, abs_tol, norm_fac
       integer(c_int) :: max_iter, precond, x0_is_input, iters
       real(c_double) :: res_start, res_final
     end function b200_pde_filter_apply

     integer(c_int) function b200_sumab(ue_d, ve_d, we_d, u_d, v_d, w_d, &
          ulag1_d, vlag1_d, wlag1_d, ulag2_d, vlag2_d, wlag2_d, ab, nab, n, &
          stream) bind(c, name='b200_sumab')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: ue_d, ve_d, we_d, u_d, v_d, w_d
       type(c_ptr), value :: ulag1_d, vlag1_d, wlag1_d, ulag2_d, vlag2_d, wlag2_d
       real(c_double), dimension(*) :: ab
       integer(c_int) :: nab, n
       type(c_ptr), value :: stream
     end function b200_sumab

     integer(c_int) function b200_makeabf(abx1_d, aby1_d, abz1_d, abx2_d, &
          aby2_d, abz2_d, fx_d, fy_d, fz_d, rho, ext, n, stream) &
          bind(c, name='b200_makeabf')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: abx1_d, aby1_d, abz1_d, abx2_d, aby2_d, abz2_d
       type(c_ptr), value :: fx_d, fy_d, fz_d
       real(c_double) :: rho
       real(c_double), dimension(*) :: ext
       integer(c_int) :: n
       type(c_ptr), value :: stream
     end function b200_makeabf

     integer(c_int) function b200_makebdf(ulag1_d, vlag1_d, wlag1_d, ulag2_d, &
          vlag2_d, wlag2_d, fx_d, fy_d, fz_d, u_d, v_d, w_d, B_d, rho, dt, bd, &
          nbd, n, stream) bind(c, name='b200_makebdf')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: ulag1_d, vlag1_d, wlag1_d, ulag2_d, vlag2_d, wlag2_d
       type(c_ptr), value :: fx_d, fy_d, fz_d, u_d, v_d, w_d, B_d
       real(c_double) :: rho, dt
       real(c_double), dimension(*) :: bd
       integer(c_int) :: nbd, n
       type(c_ptr), value :: stream
     end function b200_makebdf

     integer(c_int) function b200_makeabf_bdf(abx1_d, aby1_d, abz1_d, abx2_d, &
          aby2_d, abz2_d, ulag1_d, vlag1_d, wlag1_d, ulag2_d, vlag2_d, wlag2_d, &
          fx_d, fy_d, fz_d, u_d, v_d, w_d, B_d, rho, dt, ext, bd, nbd, n, &
          stream) bind(c, name='b200_makeabf_bdf')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: abx1_d, aby1_d, abz1_d, abx2_d, aby2_d, abz2_d
       type(c_ptr), value :: ulag1_d, vlag1_d, wlag1_d, ulag2_d, vlag2_d, wlag2_d
       type(c_ptr), value :: fx_d, fy_d, fz_d, u_d, v_d, w_d, B_d
       real(c_double) :: rho, dt
       real(c_double), dimension(*) :: ext, bd
       integer(c_int) :: nbd, n
       type(c_ptr), value :: stream
     end function b200_makeabf_bdf

     integer(c_int) function b200_adjrhs_set_element_order(handle, nelem, &
          order) bind(c, name='b200_adjrhs_set_element_order')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       integer(c_int) :: nelem
       integer(c_int), dimension(*) :: order
     end function b200_adjrhs_set_element_order

     integer(c_int) function b200_adjrhs_set_gs_fused(handle, flag) &
          bind(c, name='b200_adjrhs_set_gs_fused')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       integer(c_int) :: flag
     end function b200_adjrhs_set_gs_fused

     integer(c_int) function b200_adjrhs_gs_info(handle, fused, &
          classes_in_kernel, classes_total) &
          bind(c, name='b200_adjrhs_gs_info')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       integer(c_int) :: fused
       integer(c_int64_t) :: classes_in_kernel, classes_total
     end function b200_adjrhs_gs_info

     integer(c_int) function b200_adjrhs_set_xstage(handle, flag) &
          bind(c, name='b200_adjrhs_set_xstage')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       integer(c_int) :: flag
     end function b200_adjrhs_set_xstage

     integer(c_int) function b200_adjrhs_xstage_info(handle, active, &
          classes_staged, classes_left, classes_total) &
          bind(c, name='b200_adjrhs_xstage_info')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       integer(c_int) :: active
       integer(c_int64_t) :: classes_staged, classes_left, classes_total
     end function b200_adjrhs_xstage_info

     integer(c_int) function b200_adjrhs_set_dealias(handle, flag) &
          bind(c, name='b200_adjrhs_set_dealias')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       integer(c_int) :: flag
     end function b200_adjrhs_set_dealias

     integer(c_int) function b200_brinkman_compute(fu_d, fv_d, fw_d, u_d, &
          v_d, w_d, chi_d, n, stream) bind(c, name='b200_brinkman_compute')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: fu_d, fv_d, fw_d, u_d, v_d, w_d, chi_d
       integer(c_int) :: n
       type(c_ptr), value :: stream
     end function b200_brinkman_compute

     integer(c_int) function b200_lube_compute(fu_d, fv_d, fw_d, u_d, v_d, &
          w_d, chi_d, K, mask_d, mask_size, n, stream) &
          bind(c, name='b200_lube_compute')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: fu_d, fv_d, fw_d, u_d, v_d, w_d, chi_d
       real(c_double) :: K
       type(c_ptr), value :: mask_d
       integer(c_int) :: mask_size, n
       type(c_ptr), value :: stream
     end function b200_lube_compute

     integer(c_int) function b200_opcolv(fx_d, fy_d, fz_d, B_d, n, stream) &
          bind(c, name='b200_opcolv')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: fx_d, fy_d, fz_d, B_d
       integer(c_int) :: n
       type(c_ptr), value :: stream
     end function b200_opcolv

     integer(c_int) function b200_ramp_forward(chi_d, rho_d, n, f_min, &
          f_max, q, convex_up, stream) bind(c, name='b200_ramp_forward')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: chi_d, rho_d
       integer(c_int) :: n, convex_up
       real(c_double) :: f_min, f_max, q
       type(c_ptr), value :: stream
     end function b200_ramp_forward

     integer(c_int) function b200_ramp_backward(dF_drho_d, dF_dchi_d, rho_d, &
          n, f_min, f_max, q, convex_up, stream) &
          bind(c, name='b200_ramp_backward')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: dF_drho_d, dF_dchi_d, rho_d
       integer(c_int) :: n, convex_up
       real(c_double) :: f_min, f_max, q
       type(c_ptr), value :: stream
     end function b200_ramp_backward

     integer(c_int) function b200_sensitivity(sens_d, u_d, v_d, w_d, ua_d, &
          va_d, wa_d, K_obj, if_lube, n, stream) &
          bind(c, name='b200_sensitivity')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: sens_d, u_d, v_d, w_d, ua_d, va_d, wa_d
       real(c_double) :: K_obj
       integer(c_int) :: if_lube, n
       type(c_ptr), value :: stream
     end function b200_sensitivity

     integer(c_int) function b200_steady_field_update(res, x_d, x_old_d, n, &
          stream) bind(c, name='b200_steady_field_update')
       use, intrinsic :: iso_c_binding
       real(c_double) :: res
       type(c_ptr), value :: x_d, x_old_d
       integer(c_int) :: n
       type(c_ptr), value :: stream
     end function b200_steady_field_update

     ! ---- gather-scatter ---------------------------------------------------
     integer(c_int) function b200_gs_init(handle, key, on_device) &
          bind(c, name='b200_gs_init')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: key          ! c_loc(dof%dof) or a device pointer
       integer(c_int) :: on_device
     end function b200_gs_init

     integer(c_int) function b200_gs_get_classes(handle, class_id, nclass) &
          bind(c, name='b200_gs_get_classes')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       integer(c_int64_t) :: class_id(*), nclass
     end function b200_gs_get_classes

     integer(c_int) function b200_gs_op(handle, f_d) bind(c, name='b200_gs_op')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle, f_d
     end function b200_gs_op

     integer(c_int) function b200_gs_op3(handle, fx_d, fy_d, fz_d) &
          bind(c, name='b200_gs_op3')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle, fx_d, fy_d, fz_d
     end function b200_gs_op3

     ! ---- multi-GPU --------------------------------------------------------
     integer(c_int) function b200_comm_unique_id(id128) &
          bind(c, name='b200_comm_unique_id')
       use, intrinsic :: iso_c_binding
       character(kind=c_char) :: id128(128)
     end function b200_comm_unique_id

     integer(c_int) function b200_comm_init(handle, id128, rank, nranks) &
          bind(c, name='b200_comm_init')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       character(kind=c_char) :: id128(128)
       integer(c_int) :: rank, nranks
     end function b200_comm_init

     integer(c_int) function b200_gs_init_shared(handle, nshared, shared_dof, &
          nneigh, neigh_rank, neigh_off, neigh_idx) &
          bind(c, name='b200_gs_init_shared')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       integer(c_int) :: nshared, nneigh
       integer(c_int) :: shared_dof(*), neigh_rank(*), neigh_off(*), neigh_idx(*)
     end function b200_gs_init_shared

     integer(c_int) function b200_adjrhs_set_boundary_elements(handle, nbnd, &
          bnd_elem) bind(c, name='b200_adjrhs_set_boundary_elements')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       integer(c_int) :: nbnd
       integer(c_int) :: bnd_elem(*)
     end function b200_adjrhs_set_boundary_elements

     integer(c_int) function b200_gs_init_shared_from_keys(handle, key, &
          on_device, cand, nshared, nneigh) &
          bind(c, name='b200_gs_init_shared_from_keys')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       type(c_ptr), value :: key
       integer(c_int) :: on_device
       type(c_ptr), value :: cand
       integer(c_int) :: nshared, nneigh
     end function b200_gs_init_shared_from_keys

     ! ---- diagnostics ------------------------------------------------------
     integer(c_int) function b200_adjrhs_enable_timing(handle, flag) &
          bind(c, name='b200_adjrhs_enable_timing')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       integer(c_int) :: flag
     end function b200_adjrhs_enable_timing

     integer(c_int) function b200_adjrhs_get_timing(handle, elem_kernel_ms, &
          gs_ms, launches) bind(c, name='b200_adjrhs_get_timing')
       use, intrinsic :: iso_c_binding
       type(c_ptr), value :: handle
       real(c_double) :: elem_kernel_ms, gs_ms
       integer(c_int64_t) :: launches
     end function b200_adjrhs_get_timing
  end interface

contains

  !> Ends the job through `neko_error` when a library call returned non-zero.
  !! @param ierr   status returned by a `b200_*` function
  !! @param where  name of the entry point, for the message
  subroutine b200_check(ierr, where)
    use utils, only : neko_error
    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: where
    character(kind=c_char), dimension(:), pointer :: cmsg
    character(len=512) :: msg
    type(c_ptr) :: p
    integer :: i

    if (ierr .eq. 0) return
    msg = ''
    p = b200_last_error()
    if (c_associated(p)) then
       call c_f_pointer(p, cmsg, [512])
       do i = 1, 512
          if (cmsg(i) .eq. c_null_char) exit
          msg(i:i) = cmsg(i)
       end do
    end if
    call neko_error(where // ' failed: ' // trim(msg))
  end subroutine b200_check

end module neko_top_b200
