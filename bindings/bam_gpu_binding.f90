!~**********************************************************************
!~* bam_gpu_c: ISO_C_BINDING interface to the B200 hot path (include/bam_gpu.h).
!~*
!~* SOURCE ONLY: this image has no Fortran compiler (gfortran/flang/nvfortran absent) and BMSL/miniDMSL are
!~* not vendored, so this file has never been compiled here.  It is deliberately small and mechanical: one
!~* bind(C) interface per entry point of include/bam_gpu.h, with the C argument order and types
!~* (int32_t = integer(c_int), double = real(c_double), int64_t = integer(c_int64_t), char* = character(c_char) array).
!~* See INTEGRATION.md for the three-line patch of src/BaM_tools.f90 / src/BaM_main.f90 that uses it.
!~**********************************************************************
module bam_gpu_c
use, intrinsic :: iso_c_binding
implicit none
public

integer(c_int), parameter :: BAMGPU_MESS_LEN = 250, BAMGPU_PRIOR_NPAR = 4, BAMGPU_NENV = 7

! mirrors `struct bamgpu_problem` field by field (same order)
type, bind(C) :: bamgpu_problem
    type(c_ptr)    :: model_id                      ! NUL-terminated "BaRatin", "Linear", ...
    integer(c_int) :: nobs, nX, nY, ntheta
    type(c_ptr)    :: X, Xu, Xb, Xbindx, Y, Yu, Yb, Ybindx
    type(c_ptr)    :: partype, nval, fix_value, var_indx
    type(c_ptr)    :: prior_theta_dist, prior_theta_par
    type(c_ptr)    :: sigma_func, nremnant, prior_gamma_dist, prior_gamma_par
    type(c_ptr)    :: priorM
    type(c_ptr)    :: xtra_i
    integer(c_int) :: n_xtra_i
    type(c_ptr)    :: xtra_d
    integer(c_int) :: n_xtra_d
    type(c_ptr)    :: xtra_text
    real(c_double) :: mv, unfeas_flag
end type bamgpu_problem

type, bind(C) :: bamgpu_sizes
    integer(c_int) :: nFit, nGamma, nLatent, nXbias, nYbias, nInfer, nDpar, nCombo, nState
end type bamgpu_sizes

interface
    integer(c_int) function bamgpu_create(h, p, device, mess) bind(C, name='bamgpu_create')
        import; type(c_ptr), intent(out) :: h; type(bamgpu_problem), intent(in) :: p
        integer(c_int), value :: device; character(kind=c_char) :: mess(*)
    end function
    subroutine bamgpu_destroy(h) bind(C, name='bamgpu_destroy')
        import; type(c_ptr), value :: h
    end subroutine
    integer(c_int) function bamgpu_get_sizes(h, s) bind(C, name='bamgpu_get_sizes')
        import; type(c_ptr), value :: h; type(bamgpu_sizes), intent(out) :: s
    end function
    integer(c_int) function bamgpu_dist_id(name) bind(C, name='bamgpu_dist_id')
        import; character(kind=c_char) :: name(*)
    end function
    integer(c_int) function bamgpu_sigmafunc_id(name) bind(C, name='bamgpu_sigmafunc_id')
        import; character(kind=c_char) :: name(*)
    end function
    ! replaces Posterior_wrapper (src/BaM_tools.f90:3467): K vectors at once, x(P,K)
    integer(c_int) function bamgpu_logpost(h, K, x, fx, feas, isnull, dpar, mess) bind(C, name='bamgpu_logpost')
        import; type(c_ptr), value :: h; integer(c_int), value :: K
        real(c_double), intent(in) :: x(*); real(c_double), intent(out) :: fx(*)
        integer(c_int), intent(out) :: feas(*), isnull(*); type(c_ptr), value :: dpar
        character(kind=c_char) :: mess(*)
    end function
    ! replaces Adaptive_Metro_OAAT as called by BaM_Fit (src/BaM_tools.f90:615)
    integer(c_int) function bamgpu_fit(h, K, start, start_std, nAdapt, nCycles, MinMoveRate, MaxMoveRate, &
                                       DownMult, UpMult, seed, burn, keep_every, out_rows, n_keep, mess) &
                                       bind(C, name='bamgpu_fit')
        import; type(c_ptr), value :: h; integer(c_int), value :: K, nAdapt, nCycles, burn, keep_every
        real(c_double), intent(in) :: start(*), start_std(*)
        real(c_double), value :: MinMoveRate, MaxMoveRate, DownMult, UpMult
        integer(c_int64_t), value :: seed; real(c_double), intent(out) :: out_rows(*)
        integer(c_int64_t), intent(out) :: n_keep; character(kind=c_char) :: mess(*)
    end function
    ! replaces the replication loop of BaM_Prediction + BaM_ProcessSpag (src/BaM_tools.f90:1034-1108,1393-1426)
    integer(c_int) function bamgpu_predict(h, Nrep, mc, mode, nobs_pred, xspag, nspag, doParametric, doRemnant, &
                                           seed, t0, t1, env, spag, mess) bind(C, name='bamgpu_predict')
        import; type(c_ptr), value :: h; integer(c_int64_t), value :: Nrep, seed
        type(c_ptr), value :: mc, mode; integer(c_int), value :: nobs_pred, doParametric, t0, t1
        type(c_ptr), intent(in) :: xspag(*); integer(c_int), intent(in) :: nspag(*), doRemnant(*)
        type(c_ptr), value :: env, spag; character(kind=c_char) :: mess(*)
    end function
    ! the same with the state spaghetti (src/BaM_tools.f90:1071-1075,1121-1132): the nState state variables follow the
    ! nY outputs in env / spag
    integer(c_int) function bamgpu_predict_states(h, Nrep, mc, mode, nobs_pred, xspag, nspag, doParametric, doRemnant, &
                                                  seed, t0, t1, env, spag, mess) bind(C, name='bamgpu_predict_states')
        import; type(c_ptr), value :: h; integer(c_int64_t), value :: Nrep, seed
        type(c_ptr), value :: mc, mode; integer(c_int), value :: nobs_pred, doParametric, t0, t1
        type(c_ptr), intent(in) :: xspag(*); integer(c_int), intent(in) :: nspag(*), doRemnant(*)
        type(c_ptr), value :: env, spag; character(kind=c_char) :: mess(*)
    end function
    ! prior / likelihood / hyper split of K vectors: the loop body of BaM_computeDIC (src/BaM_tools.f90:1626-1646)
    integer(c_int) function bamgpu_logpost_parts(h, K, x, prior, lkh, hyper, feas, isnull, mess) &
                                                 bind(C, name='bamgpu_logpost_parts')
        import; type(c_ptr), value :: h; integer(c_int), value :: K
        real(c_double), intent(in) :: x(*); real(c_double), intent(out) :: prior(*), lkh(*), hyper(*)
        integer(c_int), intent(out) :: feas(*), isnull(*); character(kind=c_char) :: mess(*)
    end function
    ! replaces GetSim_calibration (src/BaM_tools.f90:3619-3680), the model run behind BaM_Residual (:794)
    integer(c_int) function bamgpu_calibration_run(h, p, x, Xtrue, Ysim, feas, mess) &
                                                   bind(C, name='bamgpu_calibration_run')
        import; type(c_ptr), value :: h; type(bamgpu_problem), intent(in) :: p
        real(c_double), intent(in) :: x(*); real(c_double), intent(out) :: Xtrue(*), Ysim(*)
        integer(c_int), intent(out) :: feas; character(kind=c_char) :: mess(*)
    end function
    ! GetSim_calibration with its `state` argument (the state columns of Results_Residuals, :793-794,860-873)
    integer(c_int) function bamgpu_calibration_run_states(h, p, x, Xtrue, Ysim, state, feas, mess) &
                                                          bind(C, name='bamgpu_calibration_run_states')
        import; type(c_ptr), value :: h; type(bamgpu_problem), intent(in) :: p
        real(c_double), intent(in) :: x(*); real(c_double), intent(out) :: Xtrue(*), Ysim(*), state(*)
        integer(c_int), intent(out) :: feas; character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_moments(h, cnt, mean, m2, dev_ptrs, mess) bind(C, name='bamgpu_moments')
        import; type(c_ptr), value :: h; real(c_double), intent(out) :: cnt(*), mean(*), m2(*)
        integer(c_int), value :: dev_ptrs; character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_rhat(K, P, cnt, mean, m2, rhat, mess) bind(C, name='bamgpu_rhat')
        import; integer(c_int), value :: K, P; real(c_double), intent(in) :: cnt(*), mean(*), m2(*)
        real(c_double), intent(out) :: rhat(*); character(kind=c_char) :: mess(*)
    end function
end interface

contains

pure function c_str(s) result(c)  ! Fortran string -> NUL-terminated C string
    character(*), intent(in) :: s
    character(kind=c_char) :: c(len_trim(s)+1)
    integer :: i
    do i = 1, len_trim(s); c(i) = s(i:i); enddo
    c(len_trim(s)+1) = c_null_char
end function

subroutine f_mess(cm, mess)  ! C message buffer -> Fortran mess
    character(kind=c_char), intent(in) :: cm(*)
    character(*), intent(out) :: mess
    integer :: i
    mess = ''
    do i = 1, min(len(mess), BAMGPU_MESS_LEN)
        if (cm(i) == c_null_char) exit
        mess(i:i) = cm(i)
    enddo
end subroutine

end module bam_gpu_c
