!> B200 implementation of Neko-TOP's `advection_adjoint_t` plug-in and of the
!! fused explicit right-hand side, on top of libneko_top_b200.so.
!!
!! Drop-in points in the reference (paths relative to sources/):
!!   * `adv_lin_b200_t` extends `advection_adjoint_t`
!!     (adjoint/advection_adjoint.f90:43-50) exactly like `adv_lin_no_dealias_t`
!!     (adjoint/adv_adjoint_no_dealias.f90:57-77) does, so
!!     `advection_adjoint_factory` (adjoint/advection_adjoint_fctry.f90:57-96)
!!     only needs one more branch (INTEGRATION.md) and the caller at
!!     adjoint/adjoint_pnpn.f90:680-682 is untouched.
!!   * `b200_fused_adjoint_rhs` replaces adjoint/adjoint_pnpn.f90:669-682
!!     (source terms, opcolv, compute_adjoint) plus the sensitivity sweep of
!!     objectives/minimum_dissipation_objective_function.f90:260-301 and, with
!!     `with_gs`, the three gs_Xh%op calls of adjoint_pnpn.f90:755-757.
!!
!! There is no host branch: the type errors out unless NEKO_BCKND_DEVICE == 1.
!! NOTE: not compiled in the build image (no Fortran compiler there).
module adv_lin_b200
  use, intrinsic :: iso_c_binding
  use num_types, only : rp
  use advection_adjoint, only : advection_adjoint_t
  use space, only : space_t, GL
  use interpolation, only : interpolator_t
  use field, only : field_t
  use coefs, only : coef_t
  use neko_config, only : NEKO_BCKND_DEVICE
  use utils, only : neko_error
  use device, only : glb_cmd_queue
  use neko_top_b200
  implicit none
  private

  public :: b200_fused_adjoint_rhs

  !> Adjoint / linearised advection on the B200 path.
  type, public, extends(advection_adjoint_t) :: adv_lin_b200_t
     !> Opaque library handle, one per `coef_t`.
     type(c_ptr) :: handle = C_NULL_PTR
     !> Device mirror of 1/jac (only needed by compute_linear).
     type(c_ptr) :: jacinv_d = C_NULL_PTR
   contains
     procedure, pass(this) :: compute_linear => linear_advection_b200
     procedure, pass(this) :: compute_adjoint => adjoint_advection_b200
     procedure, pass(this) :: init => init_b200
     procedure, pass(this) :: free => free_b200
  end type adv_lin_b200_t

  !> Dealiased variant: replaces `adv_lin_dealias_t`
  !! (adjoint/adv_adjoint_dealias.f90:56-131).  The fine Gauss-Legendre space
  !! and the GLL->GL interpolator are built exactly as `init_dealias` does
  !! (:143-146); the library interpolates the geometric factors (coef_GL,
  !! :153-161) and owns all fine-grid storage, so none of the 20 work arrays
  !! of :166-208 exist here.
  type, public, extends(adv_lin_b200_t) :: adv_lin_dealias_b200_t
     type(space_t) :: Xh_GL
     type(interpolator_t) :: GLL_to_GL
   contains
     procedure, pass(this) :: compute_linear => linear_advection_dealias_b200
     procedure, pass(this) :: compute_adjoint => adjoint_advection_dealias_b200
     procedure, pass(this) :: init_dealias => init_dealias_b200
  end type adv_lin_dealias_b200_t

contains

  !> Constructor; same arguments as `init_dealias`
  !! (adjoint/adv_adjoint_dealias.f90:137-161).
  !! @param lxd  number of Gauss-Legendre points; must be 3*lx/2
  !! @param coef The coefficients of the (space, mesh) pair.
  subroutine init_dealias_b200(this, lxd, coef)
    class(adv_lin_dealias_b200_t), intent(inout) :: this
    integer, intent(in) :: lxd
    type(coef_t), intent(inout), target :: coef
    integer(c_int) :: lxd_c

    call this%adv_lin_b200_t%init(coef)
    call this%Xh_GL%init(GL, lxd, lxd, lxd)
    call this%GLL_to_GL%init(this%Xh_GL, coef%Xh)
    lxd_c = lxd
    ! GLL_to_GL%Xh_to_Yh is the (lxd x lx) interpolation matrix J(a,l)
    call b200_check(b200_adv_dealias_init(this%handle, lxd_c, &
         this%GLL_to_GL%Xh_to_Yh, this%Xh_GL%dx, this%Xh_GL%wx), &
         'b200_adv_dealias_init')
  end subroutine init_dealias_b200

  !> Same argument list as `compute_adjoint_advection_dealias`; `f` in/out.
  subroutine adjoint_advection_dealias_b200(this, vx, vy, vz, vxb, vyb, vzb, &
       fx, fy, fz, Xh, coef, n)
    class(adv_lin_dealias_b200_t), intent(inout) :: this
    type(space_t), intent(inout) :: Xh
    type(coef_t), intent(inout) :: coef
    type(field_t), intent(inout) :: vx, vy, vz
    type(field_t), intent(inout) :: vxb, vyb, vzb
    type(field_t), intent(inout) :: fx, fy, fz
    integer, intent(in) :: n

    call b200_check(b200_adv_adjoint_dealias_compute(this%handle, vx%x_d, vy%x_d, &
         vz%x_d, vxb%x_d, vyb%x_d, vzb%x_d, fx%x_d, fy%x_d, fz%x_d), &
         'b200_adv_adjoint_dealias_compute')
  end subroutine adjoint_advection_dealias_b200

  !> Same argument list as `compute_linear_advection_dealias`; `f` in/out.
  subroutine linear_advection_dealias_b200(this, vx, vy, vz, vxb, vyb, vzb, &
       fx, fy, fz, Xh, coef, n)
    class(adv_lin_dealias_b200_t), intent(inout) :: this
    type(space_t), intent(inout) :: Xh
    type(coef_t), intent(inout) :: coef
    type(field_t), intent(inout) :: vx, vy, vz
    type(field_t), intent(inout) :: vxb, vyb, vzb
    type(field_t), intent(inout) :: fx, fy, fz
    integer, intent(in) :: n

    call b200_check(b200_adv_linear_dealias_compute(this%handle, vx%x_d, vy%x_d, &
         vz%x_d, vxb%x_d, vyb%x_d, vzb%x_d, fx%x_d, fy%x_d, fz%x_d), &
         'b200_adv_linear_dealias_compute')
  end subroutine linear_advection_dealias_b200

  !> Constructor; same argument as `init_no_dealias`.
  !! @param coef The coefficients of the (space, mesh) pair.
  subroutine init_b200(this, coef)
    class(adv_lin_b200_t), intent(inout) :: this
    type(coef_t), intent(in) :: coef
    integer(c_int) :: lx, nelv, dev

    if (NEKO_BCKND_DEVICE .ne. 1) then
       call neko_error('adv_lin_b200_t needs the CUDA device backend')
    end if
    call this%free()
    lx = coef%Xh%lx
    nelv = coef%msh%nelv
    dev = -1 ! the CUDA device Neko selected for this rank (device_init)
    call b200_check(b200_adjrhs_create(this%handle, lx, nelv, dev), &
         'b200_adjrhs_create')
    call b200_check(b200_adjrhs_set_stream(this%handle, glb_cmd_queue), &
         'b200_adjrhs_set_stream')
    call b200_check(b200_adjrhs_set_space(this%handle, coef%Xh%dx, coef%Xh%wx), &
         'b200_adjrhs_set_space')
    call b200_check(b200_adjrhs_set_geometry(this%handle, &
         coef%drdx_d, coef%dsdx_d, coef%dtdx_d, &
         coef%drdy_d, coef%dsdy_d, coef%dtdy_d, &
         coef%drdz_d, coef%dsdz_d, coef%dtdz_d, coef%B_d), &
         'b200_adjrhs_set_geometry')
    this%jacinv_d = coef%jacinv_d
  end subroutine init_b200

  !> Destructor.
  subroutine free_b200(this)
    class(adv_lin_b200_t), intent(inout) :: this

    if (c_associated(this%handle)) then
       call b200_check(b200_adjrhs_free(this%handle), &
            'b200_adjrhs_free')
    end if
    this%handle = C_NULL_PTR
  end subroutine free_b200

  !> f -= (grad U_b)^T u_adj (weak) + int grad v . (U_b (x) u_adj); `f` in/out.
  !! Same argument list as `adjoint_advection_no_dealias`.
  subroutine adjoint_advection_b200(this, vx, vy, vz, vxb, vyb, vzb, &
       fx, fy, fz, Xh, coef, n)
    class(adv_lin_b200_t), intent(inout) :: this
    type(space_t), intent(inout) :: Xh
    type(coef_t), intent(inout) :: coef
    type(field_t), intent(inout) :: vx, vy, vz
    type(field_t), intent(inout) :: vxb, vyb, vzb
    type(field_t), intent(inout) :: fx, fy, fz
    integer, intent(in) :: n

    call b200_check(b200_adv_adjoint_compute(this%handle, vx%x_d, vy%x_d, vz%x_d, &
         vxb%x_d, vyb%x_d, vzb%x_d, fx%x_d, fy%x_d, fz%x_d), &
         'b200_adv_adjoint_compute')
  end subroutine adjoint_advection_b200

  !> f -= u'.grad U_b + U_b.grad u' (B-weighted); `f` in/out.
  subroutine linear_advection_b200(this, vx, vy, vz, vxb, vyb, vzb, &
       fx, fy, fz, Xh, coef, n)
    class(adv_lin_b200_t), intent(inout) :: this
    type(space_t), intent(inout) :: Xh
    type(coef_t), intent(inout) :: coef
    type(field_t), intent(inout) :: vx, vy, vz
    type(field_t), intent(inout) :: vxb, vyb, vzb
    type(field_t), intent(inout) :: fx, fy, fz
    integer, intent(in) :: n

    call b200_check(b200_adv_linear_compute(this%handle, vx%x_d, vy%x_d, vz%x_d, &
         vxb%x_d, vyb%x_d, vzb%x_d, this%jacinv_d, fx%x_d, fy%x_d, fz%x_d), &
         'b200_adv_linear_compute')
  end subroutine linear_advection_b200

  !> The whole explicit RHS of the adjoint momentum equation in one kernel pass.
  !! Replaces, in `adjoint_pnpn_step`:
  !!   call this%source_term%compute(t, tstep)          (Brinkman + lube [+ static])
  !!   call device_opcolv(f_x%x_d, ..., c_Xh%B_d, ...)
  !!   call this%adv%compute_adjoint(u, v, w, u_b, v_b, w_b, f_x, f_y, f_z, ...)
  !! @param adv       the B200 advection object (owns the handle)
  !! @param u,v,w     adjoint velocity
  !! @param u_b..w_b  base flow (= primal velocity in the steady problem)
  !! @param rho       filtered design (RAMP is applied in-kernel); pass `chi`
  !!                  instead to use an already mapped Brinkman amplitude
  !! @param f_x..f_z  RHS, WRITE-ONLY here
  !! @param sens      optional sensitivity field dF/dchi
  !! @param fs_x..z   optional precomputed static forcing (e.g. curl curl u)
  !! @param with_gs   also apply gs_op(f, GS_OP_ADD) (needs b200_gs_init)
  subroutine b200_fused_adjoint_rhs(adv, u, v, w, u_b, v_b, w_b, f_x, f_y, f_z, &
       rho, chi, sens, fs_x, fs_y, fs_z, with_gs)
    type(adv_lin_b200_t), intent(inout) :: adv
    type(field_t), intent(inout) :: u, v, w, u_b, v_b, w_b, f_x, f_y, f_z
    type(field_t), intent(inout), optional :: rho, chi, sens, fs_x, fs_y, fs_z
    logical, intent(in), optional :: with_gs
    type(c_ptr) :: rho_d, chi_d, sens_d, fsx_d, fsy_d, fsz_d
    logical :: gs

    rho_d = C_NULL_PTR; chi_d = C_NULL_PTR; sens_d = C_NULL_PTR
    fsx_d = C_NULL_PTR; fsy_d = C_NULL_PTR; fsz_d = C_NULL_PTR
    if (present(rho)) rho_d = rho%x_d
    if (present(chi)) chi_d = chi%x_d
    if (present(sens)) sens_d = sens%x_d
    if (present(fs_x)) then
       fsx_d = fs_x%x_d; fsy_d = fs_y%x_d; fsz_d = fs_z%x_d
    end if
    gs = .false.
    if (present(with_gs)) gs = with_gs

    if (gs) then
       call b200_check(b200_adjrhs_step(adv%handle, u%x_d, v%x_d, w%x_d, &
            u_b%x_d, v_b%x_d, w_b%x_d, rho_d, chi_d, fsx_d, fsy_d, fsz_d, &
            f_x%x_d, f_y%x_d, f_z%x_d, sens_d, C_NULL_PTR), &
            'b200_adjrhs_step')
    else
       call b200_check(b200_adjrhs_compute(adv%handle, u%x_d, v%x_d, w%x_d, &
            u_b%x_d, v_b%x_d, w_b%x_d, rho_d, chi_d, fsx_d, fsy_d, fsz_d, &
            f_x%x_d, f_y%x_d, f_z%x_d, sens_d, C_NULL_PTR), &
            'b200_adjrhs_compute')
    end if
  end subroutine b200_fused_adjoint_rhs

end module adv_lin_b200
