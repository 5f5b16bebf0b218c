!!! ------------------------------------------------------------------------
!!! mcmc_gpu.F90 -- ISO_C_BINDING interface of libmcmcb200.so (include/mcmcb200.h, ABI version 3)
!!!
!!! Added to the reference's LIBSRC1 (Makefile:92-95) together with MCMC_run_gpu.F90 (#included into module
!!! mcmcmod like the other MCMC_run_*.F90 files, mcmc.F90:77-101); host/reference_gpu.patch holds the three
!!! small edits of mcmc.F90, mcmc_main.F90 and the Makefile.  Nothing else of the reference changes: namelist
!!! &mcmc, `initialize`, MCMC_writechains and the output files stay as they are.
!!!
!!! The batch fields the reference does not have come from an OPTIONAL second namelist group in the same file,
!!!   &mcmcb nchains = 1048576, ngpus = 8, gpumodel = 'expreg', datafile = 'data.dat', seed = 1 /
!!! so that existing mcmcinit.nml files still parse (nchains defaults to 1 = the reference's own CPU loop).
!!! ------------------------------------------------------------------------
module mcmc_gpu
  use, intrinsic :: iso_c_binding
  implicit none
  public

  integer(c_int), parameter :: MCMCB_ABI_VERSION = 3
  integer(c_int), parameter :: MCMCB_DRAM = 0, MCMCB_RAM = 1, MCMCB_SCAM = 2, MCMCB_ER = 3

  !! mirror of struct mcmcb_config (include/mcmcb200.h), field for field
  type, bind(c) :: mcmcb_config
     integer(c_int) :: abi_version, method, nsimu
     integer(c_int) :: doadapt, adaptint, adapthist, adaptend, initcmatn
     integer(c_int) :: doburnin, burnintime, badaptint, greedy
     real(c_double) :: scalelimit, scalefactor, drscale, condmax, N0, S02
     integer(c_int) :: updatesigma
     real(c_double) :: alphatarget, nuparam
     integer(c_long_long) :: nchains, chain_offset, seed
     integer(c_int) :: rng_mode, device, store_chains, lanes_per_chain, dump_stride, kernel
     integer(c_int) :: pool_adapt, diag_stride, diag_lags, ngpus
     character(kind=c_char) :: model(32)
  end type mcmcb_config

  !! namelist &mcmcb
  integer(c_long_long), save :: nchains = 1
  integer, save :: ngpus = 1, pool_adapt = 0, dump_stride = 0, diag_stride = 0, store_chains = 1
  integer(c_long_long), save :: seed = 0
  character(len=31), save :: gpumodel = 'expreg'
  character(len=256), save :: datafile = 'data.dat'
  namelist /mcmcb/ nchains, ngpus, pool_adapt, dump_stride, diag_stride, store_chains, seed, gpumodel, datafile

  interface
     integer(c_int) function mcmcb_default_config(cfg) bind(c, name='mcmcb_default_config')
       import :: c_int, mcmcb_config
       type(mcmcb_config), intent(out) :: cfg
     end function mcmcb_default_config
     integer(c_int) function mcmcb_create(cfg, h) bind(c, name='mcmcb_create')
       import :: c_int, c_ptr, mcmcb_config
       type(mcmcb_config), intent(in) :: cfg
       type(c_ptr), intent(out) :: h
     end function mcmcb_create
     integer(c_int) function mcmcb_destroy(h) bind(c, name='mcmcb_destroy')
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function mcmcb_destroy
     integer(c_int) function mcmcb_set_data(h, blob, n) bind(c, name='mcmcb_set_data')
       import :: c_int, c_ptr, c_double, c_size_t
       type(c_ptr), value :: h
       real(c_double), intent(in) :: blob(*)
       integer(c_size_t), value :: n
     end function mcmcb_set_data
     integer(c_int) function mcmcb_set_initial(h, npar, nycol, par0, stride, cmat0, sigma2, nobs) bind(c, name='mcmcb_set_initial')
       import :: c_int, c_ptr, c_double, c_long_long
       type(c_ptr), value :: h
       integer(c_int), value :: npar, nycol
       real(c_double), intent(in) :: par0(*), cmat0(*), sigma2(*)
       integer(c_long_long), value :: stride
       integer(c_int), intent(in) :: nobs(*)
     end function mcmcb_set_initial
     integer(c_int) function mcmcb_run(h, nsteps) bind(c, name='mcmcb_run')
       import :: c_int, c_ptr
       type(c_ptr), value :: h
       integer(c_int), value :: nsteps
     end function mcmcb_run
     integer(c_int) function mcmcb_sync(h) bind(c, name='mcmcb_sync')
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function mcmcb_sync
     integer(c_int) function mcmcb_fetch_chain(h, ichain, ld, chain, sschain, s2chain, nrows) bind(c, name='mcmcb_fetch_chain')
       import :: c_int, c_ptr, c_double, c_long_long
       type(c_ptr), value :: h
       integer(c_long_long), value :: ichain
       integer(c_int), value :: ld
       real(c_double), intent(out) :: chain(ld,*), sschain(ld,*), s2chain(ld,*)
       integer(c_int), intent(out) :: nrows
     end function mcmcb_fetch_chain
     integer(c_int) function mcmcb_fetch_stats(h, ichain, mean, cmat, wsum, R, sigma2, counters) bind(c, name='mcmcb_fetch_stats')
       import :: c_int, c_ptr, c_double, c_long_long
       type(c_ptr), value :: h
       integer(c_long_long), value :: ichain
       real(c_double), intent(out) :: mean(*), cmat(*), wsum, R(*), sigma2(*)
       integer(c_long_long), intent(out) :: counters(8)
     end function mcmcb_fetch_stats
  end interface

contains

  !! read the optional group &mcmcb from the namelist file; absent group or file = defaults (CPU path)
  subroutine read_mcmcb_namelist(nmlfile)
    character(len=*), intent(in) :: nmlfile
    integer :: fstat
    open(unit=11, file=nmlfile, status='old', iostat=fstat)
    if (fstat /= 0) return
    read(11, nml=mcmcb, iostat=fstat)
    close(11)
  end subroutine read_mcmcb_namelist

  !! Fortran string -> NUL terminated C char array
  subroutine to_cstring(s, c)
    character(len=*), intent(in) :: s
    character(kind=c_char), intent(out) :: c(:)
    integer :: i, n
    c = c_null_char
    n = min(len_trim(s), size(c) - 1)
    do i = 1, n
       c(i) = s(i:i)
    end do
  end subroutine to_cstring

end module mcmc_gpu
