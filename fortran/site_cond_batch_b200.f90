!===============================================================================
! site_cond_batch_b200.f90 -- BASELINE configs[0]: square site percolation, L = 100,
! p = 0.60, 1000 realizations: labeling + spanning test + conductance.  The trial
! loop `do ii = 1, numtrials` of the reference (Fortran/Square/site_perc.f:87,
! bond_cond.f:123) becomes ONE call: every realization is generated, labeled and
! solved on the device (one CTA per realization runs the whole linbcg loop).
! Not compile-tested here (no Fortran compiler in the image).
!===============================================================================
program site_cond_batch_b200
  use iso_c_binding
  use perc_iface
  implicit none
  integer(c_int32_t), parameter :: m = 100, n = 100, pbc = 0, numtrials = 1000
  integer(c_int64_t) :: h, stats(16)
  integer(c_int32_t) :: rc, ii, tsites, iters(numtrials)
  real(c_double) :: G(2, numtrials), ps

  ps = 0.60d0
  tsites = int(ps * dble(m * n))                ! Sq/site.f:164
  rc = perc_create(h, PERC_SQUARE, m, n, pbc, 0)
  if (rc /= 0) stop 'perc_create'
  ! reference solver settings (Sq/bondc.f:545): tol = 1e-8, itmax = 2500, leak 1e-12, read-out threshold 1e-10
  rc = perc_batch_conduct(h, PERC_SITE, numtrials, 58302_c_int64_t, 0_c_int64_t, tsites, 0, &
                          1.0d0, 1.0d0, 1.0d-12, 1.0d-8, 2500, 1.0d-10, G, iters, stats)
  if (rc /= 0) stop 'perc_batch_conduct'
  open (unit=11, file='sitecond.txt', status='unknown')
  do ii = 1, numtrials                          ! trial, Gbot, Gtop, avg  (format of Sq/bond_cond.f:481-482,505)
    if (iters(ii) >= 0) write (11, '(i10,",",f12.9,",",f12.9,",",f12.9)') ii, G(2, ii), G(1, ii), 0.5d0 * (G(1, ii) + G(2, ii))
  end do
  close (11)
  write (6, *) 'realizations', stats(1), ' spanning', stats(4)
  rc = perc_destroy(h)
end program
