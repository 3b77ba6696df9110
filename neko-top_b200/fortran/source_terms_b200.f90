!> B200 versions of the two point-wise source terms on the adjoint-RHS path and
!! of the steady-state simulation component's field update.  Public types keep
!! the reference's names, components and type-bound procedures
!! (source_terms/simple_brinkman_source_term.f90:52-67,
!!  simulation_components/steady_simcomp.f90:49-68); only the bodies of
!! `compute_` change, so problems/steady_state_problem.f90:147-163,364-375
!! compiles unchanged.  Shown here as the replacement bodies.
!! NOTE: not compiled in the build image (no Fortran compiler there).
module source_terms_b200
  use, intrinsic :: iso_c_binding
  use num_types, only : rp
  use field, only : field_t
  use device, only : glb_cmd_queue
  use neko_top_b200
  implicit none
  private
  public :: brinkman_compute_b200, lube_compute_b200, steady_update_b200

contains

  !> Body of `simple_brinkman_source_term_compute` (:139-153): f_i -= chi u_i.
  subroutine brinkman_compute_b200(fu, fv, fw, u, v, w, chi)
    type(field_t), intent(inout) :: fu, fv, fw
    type(field_t), intent(in) :: u, v, w, chi
    integer(c_int) :: n

    n = fu%dof%size()
    call b200_check(b200_brinkman_compute(fu%x_d, fv%x_d, fw%x_d, u%x_d, v%x_d, w%x_d, &
         chi%x_d, n, glb_cmd_queue), &
         'b200_brinkman_compute')
  end subroutine brinkman_compute_b200

  !> Body of `adjoint_lube_source_term_compute`
  !! (source_terms/adjoint_lube_source_term.f90:173-206): f_i += K chi u_i,
  !! restricted to the 1-based point-zone mask when `mask_size > 0`.
  subroutine lube_compute_b200(fu, fv, fw, u, v, w, chi, K, mask_d, mask_size)
    type(field_t), intent(inout) :: fu, fv, fw
    type(field_t), intent(in) :: u, v, w, chi
    real(kind=rp), intent(in) :: K
    type(c_ptr), intent(in) :: mask_d
    integer, intent(in) :: mask_size
    integer(c_int) :: n, ms
    real(c_double) :: Kc

    n = fu%dof%size()
    ms = mask_size
    Kc = K
    call b200_check(b200_lube_compute(fu%x_d, fv%x_d, fw%x_d, u%x_d, v%x_d, w%x_d, &
         chi%x_d, Kc, mask_d, ms, n, glb_cmd_queue), &
         'b200_lube_compute')
  end subroutine lube_compute_b200

  !> One field of `steady_simcomp_compute` (:158-176): returns the local part of
  !! glsc2(old - new, old - new) and overwrites `x_old` with `x`, in one pass.
  !! The caller still does the MPI_Allreduce that field_glsc2 would do.
  function steady_update_b200(x, x_old) result(res)
    type(field_t), intent(in) :: x
    type(field_t), intent(inout) :: x_old
    real(kind=rp) :: res
    real(c_double) :: r
    integer(c_int) :: n

    n = x%dof%size()
    call b200_check(b200_steady_field_update(r, x%x_d, x_old%x_d, n, glb_cmd_queue), &
         'b200_steady_field_update')
    res = r
  end function steady_update_b200

end module source_terms_b200
