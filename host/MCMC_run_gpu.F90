!!! ------------------------------------------------------------------------
!!! MCMC_run_gpu.F90 -- the sampling loop handed to the GPU library (libmcmcb200.so)
!!!
!!! #included into module mcmcmod (mcmc.F90:77-101) next to MCMC_run.F90 / MCMC_run_ram.F90 / MCMC_run_scam.F90 /
!!! MCMC_run_er.F90, whose place it takes when namelist group &mcmcb asks for more than one chain: it sees the
!!! module's private state (par0, cmat0, nobs, chaincmat, ...), runs `nchains` chains on `ngpus` GPUs through the
!!! C ABI and leaves chain / sschain / s2chain / chainind / simuind / chaincmat / chainmean / chainwsum / sigma2 and
!!! the counters of chain 1 exactly where MCMC_savechain and MCMC_adapt would have left them, so that
!!! MCMC_writechains (MCMC_aux.F90:17-85) writes the usual files.
!!! ------------------------------------------------------------------------
subroutine MCMC_run_gpu()
  use mcmc_gpu
  implicit none

  type(mcmcb_config) :: cfg
  type(c_ptr) :: h
  integer(c_int) :: stat, nrows
  integer(c_long_long) :: cnt(8)
  real(kind=dbl), pointer :: xy(:,:)
  real(kind=dbl), allocatable :: blob(:), Rfull(:), s2tmp(:,:)
  integer :: n, npad, fstat

  if (inited /= 1) call doerror('we have not inited')

  stat = mcmcb_default_config(cfg)
  select case (trim(method))            ! mcmc_main.F90:29-37
  case ('ram');  cfg%method = MCMCB_RAM
  case ('scam'); cfg%method = MCMCB_SCAM
  case ('er');   cfg%method = MCMCB_ER
  case default;  cfg%method = MCMCB_DRAM
  end select
  !! namelist &mcmc, 1:1 (mcmcinit.F90:74-82)
  cfg%nsimu = nsimu;        cfg%doadapt = doadapt;       cfg%adaptint = adaptint;   cfg%adapthist = adapthist
  cfg%adaptend = adaptend;  cfg%initcmatn = initcmatn;   cfg%doburnin = doburnin;   cfg%burnintime = burnintime
  cfg%badaptint = badaptint; cfg%greedy = greedy;        cfg%scalelimit = scalelimit; cfg%scalefactor = scalefactor
  cfg%drscale = drscale;    cfg%condmax = condmax;       cfg%N0 = N0;               cfg%S02 = S02
  cfg%updatesigma = updatesigma; cfg%alphatarget = alphatarget; cfg%nuparam = nuparam
  !! namelist &mcmcb
  cfg%nchains = nchains;    cfg%ngpus = ngpus;           cfg%seed = seed;           cfg%pool_adapt = pool_adapt
  cfg%dump_stride = dump_stride; cfg%diag_stride = diag_stride; cfg%store_chains = store_chains
  call to_cstring(gpumodel, cfg%model)

  stat = mcmcb_create(cfg, h)
  if (stat /= 0) then
     write(*,*) 'ERROR: mcmcb_create failed, status = ', stat
     stop
  end if

  !! what the user's ssfunction loads on its first call (testcases/mcmcrun.F90:69-86): packed for the device model
  !! `expreg` as [n, +-max|x|, x(npad), y(npad)] (mcmcf90_b200/csrc/models.cuh); other models pack their own blob
  call loaddata(datafile, xy, fstat)
  if (fstat /= 0) call doerror('could not read the data file of the GPU model')
  n = size(xy, 1)
  npad = n + mod(n, 2)
  allocate(blob(2 + 2*npad))
  blob = 0.0_dbl
  blob(1) = dble(n)
  blob(2) = maxval(abs(xy(:,1)))
  if (minval(xy(:,1)) < 0.0_dbl) blob(2) = -blob(2)
  blob(3:2+n) = xy(:,1)
  blob(3+npad:2+npad+n) = xy(:,2)
  stat = mcmcb_set_data(h, blob, int(size(blob), c_size_t))

  !! what `initialize` returned (MCMC_init.F90:45-71): every chain starts from par0
  stat = mcmcb_set_initial(h, int(npar, c_int), int(nycol, c_int), par0, 0_c_long_long, cmat0, sigma2, nobs)
  if (stat /= 0) then
     write(*,*) 'ERROR: mcmcb_set_initial failed, status = ', stat
     stop
  end if

  stat = mcmcb_run(h, int(nsimu - 1, c_int))
  if (stat == 0) stat = mcmcb_sync(h)
  if (stat /= 0) then
     write(*,*) 'ERROR: mcmcb_run failed, status = ', stat
     stop
  end if

  !! chain 1 back into the module's arrays, in MCMC_savechain's layout (MCMC_aux.F90:166-185)
  if (updatesigma /= 0) then
     stat = mcmcb_fetch_chain(h, 0_c_long_long, int(nsimu, c_int), chain, sschain, s2chain, nrows)
  else                                   ! s2chain is not allocated then (MCMC_init.F90:119-122)
     allocate(s2tmp(nsimu, nycol))
     stat = mcmcb_fetch_chain(h, 0_c_long_long, int(nsimu, c_int), chain, sschain, s2tmp, nrows)
     deallocate(s2tmp)
  end if
  chainind = nrows
  simuind  = nsimu
  allocate(Rfull(npar*npar))
  stat = mcmcb_fetch_stats(h, 0_c_long_long, chainmean, chaincmat, chainwsum, Rfull, sigma2, cnt)
  R = reshape(Rfull, (/npar, npar/))
  stayed = int(cnt(1)); bndstayed = int(cnt(2)); draccepted = int(cnt(3)); drtries = int(cnt(4))
  if (verbosity>0) write(*,*) 'note: GPU run of ', nchains, ' chains done, chain 1 stayed % = ', &
       real(stayed)/real(simuind)*100.0

  stat = mcmcb_destroy(h)
  deallocate(blob, Rfull, xy)

end subroutine MCMC_run_gpu
