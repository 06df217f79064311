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
    integer(c_int) function bamgpu_moments_window(h, first, every, mess) bind(C, name='bamgpu_moments_window')
        import; type(c_ptr), value :: h; integer(c_int64_t), value :: first; integer(c_int), value :: every
        character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_rhat(K, P, cnt, mean, m2, rhat, mess) bind(C, name='bamgpu_rhat')
        import; integer(c_int), value :: K, P; real(c_double), intent(in) :: cnt(*), mean(*), m2(*)
        real(c_double), intent(out) :: rhat(*); character(kind=c_char) :: mess(*)
    end function
    ! ---- the remaining entry points of include/bam_gpu.h, in header order -------------------------------------------
    integer(c_int) function bamgpu_set_chain_offset(h, offset) bind(C, name='bamgpu_set_chain_offset')
        import; type(c_ptr), value :: h; integer(c_int64_t), value :: offset
    end function
    integer(c_int) function bamgpu_set_option(h, name, val) bind(C, name='bamgpu_set_option')
        import; type(c_ptr), value :: h; character(kind=c_char) :: name(*); integer(c_int), value :: val
    end function
    integer(c_int) function bamgpu_dist_npar(dist) bind(C, name='bamgpu_dist_npar')
        import; integer(c_int), value :: dist
    end function
    integer(c_int) function bamgpu_sigmafunc_npar(f) bind(C, name='bamgpu_sigmafunc_npar')
        import; integer(c_int), value :: f
    end function
    integer(c_int) function bamgpu_model_id(name) bind(C, name='bamgpu_model_id')
        import; character(kind=c_char) :: name(*)
    end function
    integer(c_int) function bamgpu_prior_corr_to_M(k, C, M, mess) bind(C, name='bamgpu_prior_corr_to_M')
        import; integer(c_int), value :: k; real(c_double), intent(in) :: C(*); real(c_double), intent(out) :: M(*)
        character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_chains_upload(h, K, x, mess) bind(C, name='bamgpu_chains_upload')
        import; type(c_ptr), value :: h; integer(c_int), value :: K; real(c_double), intent(in) :: x(*)
        character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_logpost_resident(h, mess) bind(C, name='bamgpu_logpost_resident')
        import; type(c_ptr), value :: h; character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_chains_download(h, fx, feas, isnull, dpar, mess) bind(C, name='bamgpu_chains_download')
        import; type(c_ptr), value :: h; real(c_double), intent(out) :: fx(*); integer(c_int), intent(out) :: feas(*), isnull(*)
        type(c_ptr), value :: dpar; character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_sync(h, mess) bind(C, name='bamgpu_sync')
        import; type(c_ptr), value :: h; character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_predict_upload(h, Nrep, mc, mode, nobs_pred, xspag, nspag, mess) bind(C, name='bamgpu_predict_upload')
        import; type(c_ptr), value :: h; integer(c_int64_t), value :: Nrep; type(c_ptr), value :: mc, mode
        integer(c_int), value :: nobs_pred; type(c_ptr), intent(in) :: xspag(*); integer(c_int), intent(in) :: nspag(*)
        character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_predict_resident(h, doParametric, doRemnant, seed, t0, t1, mess) bind(C, name='bamgpu_predict_resident')
        import; type(c_ptr), value :: h; integer(c_int), value :: doParametric, t0, t1; integer(c_int), intent(in) :: doRemnant(*)
        integer(c_int64_t), value :: seed; character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_predict_download(h, env, mess) bind(C, name='bamgpu_predict_download')
        import; type(c_ptr), value :: h; real(c_double), intent(out) :: env(*); character(kind=c_char) :: mess(*)
    end function
    ! multi-GPU: one communicator rank per GPU (MPI hosts broadcast the 128-byte id with MPI_Bcast)
    integer(c_int) function bamgpu_comm_unique_id(id128, mess) bind(C, name='bamgpu_comm_unique_id')
        import; character(kind=c_char) :: id128(128), mess(*)
    end function
    integer(c_int) function bamgpu_comm_init_rank(c, nranks, rank, id128, device, mess) bind(C, name='bamgpu_comm_init_rank')
        import; type(c_ptr), intent(out) :: c; integer(c_int), value :: nranks, rank, device
        character(kind=c_char), intent(in) :: id128(128); character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_comm_init_all(comms, ndev, devices, mess) bind(C, name='bamgpu_comm_init_all')
        import; type(c_ptr), intent(out) :: comms(*); integer(c_int), value :: ndev; type(c_ptr), value :: devices
        character(kind=c_char) :: mess(*)
    end function
    subroutine bamgpu_comm_destroy(c) bind(C, name='bamgpu_comm_destroy')
        import; type(c_ptr), value :: c
    end subroutine
    integer(c_int) function bamgpu_comm_info(c, nranks, rank, nccl_version) bind(C, name='bamgpu_comm_info')
        import; type(c_ptr), value :: c; integer(c_int), intent(out) :: nranks, rank, nccl_version
    end function
    integer(c_int) function bamgpu_moments_allgather(h, c, rhat, count_all, mean_all, m2_all, ms, mess) &
                                                     bind(C, name='bamgpu_moments_allgather')
        import; type(c_ptr), value :: h, c; real(c_double), intent(out) :: rhat(*); type(c_ptr), value :: count_all, mean_all, m2_all, ms
        character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_envelope_allgather(h, c, nobs_pred, env_all, ms, mess) bind(C, name='bamgpu_envelope_allgather')
        import; type(c_ptr), value :: h, c; integer(c_int), value :: nobs_pred; real(c_double), intent(out) :: env_all(*)
        type(c_ptr), value :: ms; character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_comm_allreduce_sum(c, dev_buf, count, dtype, cuda_stream, mess) &
                                                      bind(C, name='bamgpu_comm_allreduce_sum')
        import; type(c_ptr), value :: c, dev_buf, cuda_stream; integer(c_size_t), value :: count; integer(c_int), value :: dtype
        character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_predict_sample_sharded(h, c, rep_base, nrep_total, doParametric, doRemnant, seed, t0, t1, mess) &
                                                          bind(C, name='bamgpu_predict_sample_sharded')
        import; type(c_ptr), value :: h, c; integer(c_int64_t), value :: rep_base, nrep_total, seed
        integer(c_int), value :: doParametric, t0, t1; integer(c_int), intent(in) :: doRemnant(*); character(kind=c_char) :: mess(*)
    end function
    ! measurement helpers
    integer(c_int64_t) function bamgpu_launch_count(h) bind(C, name='bamgpu_launch_count')
        import; type(c_ptr), value :: h
    end function
    real(c_double) function bamgpu_last_kernel_ms(h) bind(C, name='bamgpu_last_kernel_ms')
        import; type(c_ptr), value :: h
    end function
    integer(c_int) function bamgpu_step_begin(h) bind(C, name='bamgpu_step_begin')
        import; type(c_ptr), value :: h
    end function
    integer(c_int) function bamgpu_step_end(h) bind(C, name='bamgpu_step_end')
        import; type(c_ptr), value :: h
    end function
    integer(c_int) function bamgpu_step_times(h, ms, cap) bind(C, name='bamgpu_step_times')
        import; type(c_ptr), value :: h; real(c_double), intent(out) :: ms(*); integer(c_int), value :: cap
    end function
    integer(c_int) function bamgpu_kernel_times(h, ms, cap) bind(C, name='bamgpu_kernel_times')
        import; type(c_ptr), value :: h; real(c_double), intent(out) :: ms(*); integer(c_int), value :: cap
    end function
    integer(c_int) function bamgpu_phase_times(h, prep_ms, final_ms, cap) bind(C, name='bamgpu_phase_times')
        import; type(c_ptr), value :: h; real(c_double), intent(out) :: prep_ms(*), final_ms(*); integer(c_int), value :: cap
    end function
    integer(c_int) function bamgpu_flush_l2(h, bytes) bind(C, name='bamgpu_flush_l2')
        import; type(c_ptr), value :: h; integer(c_size_t), value :: bytes
    end function
    integer(c_int) function bamgpu_fp64_peak(device, tflops, mess) bind(C, name='bamgpu_fp64_peak')
        import; integer(c_int), value :: device; real(c_double), intent(out) :: tflops; character(kind=c_char) :: mess(*)
    end function
    integer(c_int) function bamgpu_pin(ptr, bytes) bind(C, name='bamgpu_pin')
        import; type(c_ptr), value :: ptr; integer(c_size_t), value :: bytes
    end function
    integer(c_int) function bamgpu_unpin(ptr) bind(C, name='bamgpu_unpin')
        import; type(c_ptr), value :: ptr
    end function
    type(c_ptr) function bamgpu_version() bind(C, name='bamgpu_version')
        import
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
