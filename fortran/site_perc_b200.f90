!===============================================================================
! site_perc_b200.f90 -- the reference's PROGRAM site_perc (Fortran/Square/site_perc.f)
! on top of libperc_b200: same parameter block, same seed fan-out (srand(58302),
! tseed(i) = int(rand(0)*1000000)+1, :69-75), same Fisher-Yates shuffle (:105-120),
! same output row "tseed, f, maxcs, perccls" (format 111, :257-260).  Only the
! fill-until-spanning loop (:133-254) is replaced: perc_first_span finds the first
! spanning step by bisection over the fill count on the GPU (spanning is monotone in
! the number of occupied sites; the partition at a given count is order independent).
! Not compile-tested in this image (no Fortran compiler).
!===============================================================================
program site_perc_b200
  use iso_c_binding
  use perc_iface
  implicit none
  integer(c_int32_t) :: m, n, t, pbc, lattice, device, rc
  integer(c_int32_t) :: numtrials, seed, i, ii, j, temp
  integer(c_int32_t) :: kstar, maxcs, perccls
  integer(c_int32_t), allocatable :: order(:), tseed(:)
  integer(c_int64_t) :: h
  real(c_float) :: f
  real :: rand

  open(unit=10, file='site_perc.txt')
  m = 50; n = 50; t = m*n; pbc = 0            ! Sq/site_perc.f:47-55
  lattice = PERC_SQUARE; device = 0
  numtrials = 1000; seed = 58302               ! :69-70
  allocate(order(t), tseed(50000))
  call srand(seed)
  do i = 1, 50000
     tseed(i) = int(rand(0)*1000000)+1         ! :73-75
  end do

  rc = perc_create(h, lattice, m, n, pbc, device)
  if (rc /= 0) stop 'perc_create failed (no CUDA device? there is no CPU fallback)'

  do ii = 1, numtrials
     call srand(tseed(ii))                      ! :105
     do i = 1, t
        order(i) = i
     end do
     do i = 1, t                                ! :115-120
        j = i + (t-i+1)*rand(0)
        temp = order(i); order(i) = order(j); order(j) = temp
     end do
     rc = perc_set_site_order(h, order)
     rc = perc_first_span(h, PERC_SITE, PERC_SITE, kstar, f, maxcs, perccls)
     write(6,*) tseed(ii), f, maxcs, perccls
     write(10,111) tseed(ii), f, maxcs, perccls
  end do
111 format(i10,",",f12.9,",",i10,",",i10)
  rc = perc_destroy(h)
end program site_perc_b200
