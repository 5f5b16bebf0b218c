!!! injrand -- replaces the compiler runtime's random_number in the PATCHED COPY of the reference that
!!! oracle/Makefile.ref builds (test infrastructure; not part of the product, not part of the reference).
!!! The patched call sites (mcmcrand.F90:55,104,138,156,177; MCMC_DRAM.F90:132,151) read the next numbers of
!!! the file uniforms.bin (raw little-endian float64) -- the same stream mcmcb_inject_uniforms gives the GPU
!!! and orc_set_rng_injected gives the C oracle.
module injrand
  implicit none
  private
  integer, parameter :: dp = selected_real_kind(15)
  real(dp), allocatable, save :: u(:)
  integer, save :: pos = 0, n = -1
  public :: inj_random_number
  interface inj_random_number
     module procedure inj_scalar, inj_vector
  end interface
contains
  subroutine load()
    integer :: sz
    if (n >= 0) return
    inquire(file='uniforms.bin', size=sz)
    n = sz/8
    allocate(u(n))
    open(977, file='uniforms.bin', access='stream', form='unformatted', status='old')
    read(977) u
    close(977)
  end subroutine load
  subroutine inj_scalar(x)
    real(dp), intent(out) :: x
    call load()
    if (pos + 1 > n) stop 'injrand: stream exhausted'
    pos = pos + 1
    x = u(pos)
  end subroutine inj_scalar
  subroutine inj_vector(x)
    real(dp), intent(out) :: x(:)
    integer :: k
    call load()
    k = size(x)
    if (pos + k > n) stop 'injrand: stream exhausted'
    x = u(pos+1:pos+k)
    pos = pos + k
  end subroutine inj_vector
end module injrand
