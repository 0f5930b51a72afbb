!===============================================================================
! site_b200.f90 -- the reference's PROGRAM site (Fortran/Square/site.f, Fortran/Triangular/site.f)
! on top of libperc_b200: same parameter block (:51-67), seed 1080115 and Fisher-Yates shuffle
! (:131-147), same output files: bondlist.txt (b1,b2  :122-124), siteorder.txt (:151-153) and
! site.txt (j, s(j), c(j)  :354-358).  The fill loop with its O(t) relabel scans (:162-289) and the
! spanning scan (:309-344) are ONE call, perc_site.  Labels are canonical (smallest member site),
! c() is indexed by them.  Not compile-tested in this image (no Fortran compiler).
!===============================================================================
program site_b200
  use iso_c_binding
  use perc_iface
  implicit none
  integer(c_int32_t) :: m, n, t, pbc, lattice, nb, rc, seed, tsites, i, j, temp
  integer(c_int32_t) :: maxcs, perccln, perccls
  integer(c_int32_t), allocatable :: order(:), s(:), c(:), blist(:)
  integer(c_int64_t) :: h
  double precision :: ps
  real :: rand

  open(unit=10, file='site.txt')
  open(unit=12, file='siteorder.txt')
  open(unit=13, file='bondlist.txt')
  m = 50; n = 50; t = m*n; pbc = 0; ps = 0.60d+00       ! Sq/site.f:51-61
  lattice = PERC_SQUARE                                 ! PERC_TRIANGULAR for Triangular/site.f (m even)
  rc = perc_geom_nb(lattice, m, n, pbc, nb)              ! nb formulas :89-93
  allocate(order(t), s(t), c(t), blist(2*nb))
  rc = perc_geom_bondlist(lattice, m, n, pbc, blist)     ! :106-120
  do i = 1, nb
     write(13,121) blist(i), blist(nb+i)
  end do

  seed = 1080115                                         ! :131
  call srand(seed)
  do i = 1, t
     order(i) = i
  end do
  do i = 1, t                                            ! :142-147
     j = i + (t-i+1)*rand(0)
     temp = order(i); order(i) = order(j); order(j) = temp
  end do
  do i = 1, t
     write(12,*) order(i)
  end do

  tsites = ps*t                                          ! :164 (fp64 product, truncated)
  rc = perc_create(h, lattice, m, n, pbc, 0)
  if (rc /= 0) stop 'perc_create failed (no CUDA device? there is no CPU fallback)'
  rc = perc_site(h, order, tsites, s, c, maxcs, perccln, perccls)
  if (rc /= 0) stop 'perc_site failed'
  write(6,*) 'largest cluster', maxcs, ' spanning cluster', perccln, ' size', perccls
  do j = 1, t
     write(10,111) j, s(j), c(j)                         ! :354-356
  end do
111 format(i10,",",i10,",",i10)
121 format(i10,",",i10)
  rc = perc_destroy(h)
end program site_b200
