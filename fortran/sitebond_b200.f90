!===============================================================================
! sitebond_b200.f90 -- the reference's PROGRAMs sitebond and bondsite (Fortran/Square/sitebond.f,
! bondsite.f and the Triangular twins): mixed site/bond percolation, int(ps*t) sites and int(pb*nb)
! bonds occupied; cluster size = sites + bonds.  Seeds sseed = 143285, bseed = 43716 (Sq/sitebond.f:66-67),
! shuffles :117-176, outputs sbsite.txt (i, s(i), c(i)) and sbbond.txt (b1, b2, label) (:469-477).
! Both fill orders (sites first / bonds first) give the same partition, so one call serves both programs:
! perc_sitebond replaces Sq/sitebond.f:187-458 and Sq/bondsite.f:182-412.
! Not compile-tested in this image (no Fortran compiler).
!===============================================================================
program sitebond_b200
  use iso_c_binding
  use perc_iface
  implicit none
  integer(c_int32_t) :: m, n, t, pbc, lattice, nb, rc, sseed, bseed, tsites, tbonds, i, j, t1, t2
  integer(c_int32_t) :: maxcs, perccln, perccls
  integer(c_int32_t), allocatable :: b(:), border(:), sorder(:), s(:), b3(:), c(:)
  integer(c_int64_t) :: h
  double precision :: ps, pb
  real :: rand

  open(unit=10, file='sbsite.txt')
  open(unit=11, file='sbbond.txt')
  m = 50; n = 50; t = m*n; pbc = 0; ps = 0.50d+00; pb = 0.50d+00     ! Sq/sitebond.f:54-61
  sseed = 143285; bseed = 43716                                     ! :66-67
  lattice = PERC_SQUARE
  rc = perc_geom_nb(lattice, m, n, pbc, nb)
  allocate(b(2*nb), border(2*nb), sorder(t), s(t), b3(nb), c(t))
  rc = perc_geom_bondlist(lattice, m, n, pbc, b)
  border = b
  call srand(sseed)                                                  ! :117-132
  do i = 1, t
     sorder(i) = i
  end do
  do i = 1, t
     j = i + (t-i+1)*rand(0)
     t1 = sorder(i); sorder(i) = sorder(j); sorder(j) = t1
  end do
  call srand(bseed)                                                  ! :165-176
  do i = 1, nb
     j = i + (nb-i+1)*rand(0)
     t1 = border(i);    border(i) = border(j);       border(j) = t1
     t2 = border(nb+i); border(nb+i) = border(nb+j); border(nb+j) = t2
  end do
  tsites = ps*t
  tbonds = pb*nb
  rc = perc_create(h, lattice, m, n, pbc, 0)
  if (rc /= 0) stop 'perc_create failed (no CUDA device? there is no CPU fallback)'
  rc = perc_sitebond(h, sorder, tsites, border, tbonds, s, b3, c, maxcs, perccln, perccls)
  if (rc /= 0) stop 'perc_sitebond failed'
  do i = 1, t
     write(10,111) i, s(i), c(i)                                     ! :469-471
  end do
  do i = 1, nb
     write(11,111) b(i), b(nb+i), b3(i)                              ! :473-475 (a bond with no occupied end: label t + i)
  end do
111 format(i10,",",i10,",",i10)
  rc = perc_destroy(h)
end program sitebond_b200
