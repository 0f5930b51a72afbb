!===============================================================================
! bond_perc_b200.f90 -- the reference's PROGRAM bond_perc (Fortran/Square/bond_perc.f): numtrials bond
! trials, each filled until the first spanning cluster; output row "tseed, fb, maxcs, perccls"
! (format 111, :364-367).  Seed fan-out srand(58302), tseed(i) = int(rand(0)*1000000)+1 (:68-74) and the
! bond shuffle are the reference's; the fill-until-spanning loop (:133-359, a full bond-list scan per
! added bond) is perc_first_span: bisection over the fill count on the GPU.
! Not compile-tested in this image (no Fortran compiler).
!===============================================================================
program bond_perc_b200
  use iso_c_binding
  use perc_iface
  implicit none
  integer(c_int32_t) :: m, n, pbc, lattice, nb, rc, numtrials, seed, i, ii, j, t1, t2
  integer(c_int32_t) :: kstar, maxcs, perccls
  integer(c_int32_t), allocatable :: b(:), border(:), tseed(:)
  integer(c_int64_t) :: h
  real(c_float) :: fb
  real :: rand

  open(unit=10, file='bond_perc.txt')
  m = 50; n = 50; pbc = 0; lattice = PERC_SQUARE
  numtrials = 10; seed = 58302                           ! Sq/bond_perc.f:68-69
  rc = perc_geom_nb(lattice, m, n, pbc, nb)
  allocate(b(2*nb), border(2*nb), tseed(50000))
  rc = perc_geom_bondlist(lattice, m, n, pbc, b)
  call srand(seed)
  do i = 1, 50000
     tseed(i) = int(rand(0)*1000000)+1
  end do
  rc = perc_create(h, lattice, m, n, pbc, 0)
  if (rc /= 0) stop 'perc_create failed (no CUDA device? there is no CPU fallback)'
  do ii = 1, numtrials
     border = b
     call srand(tseed(ii))
     do i = 1, nb
        j = i + (nb-i+1)*rand(0)
        t1 = border(i);    border(i) = border(j);       border(j) = t1
        t2 = border(nb+i); border(nb+i) = border(nb+j); border(nb+j) = t2
     end do
     rc = perc_set_bond_order(h, border)
     rc = perc_first_span(h, PERC_BOND, PERC_BOND, kstar, fb, maxcs, perccls)
     write(6,*) tseed(ii), fb, maxcs, perccls
     write(10,111) tseed(ii), fb, maxcs, perccls
  end do
111 format(i10,",",f12.9,",",i10,",",i10)
  rc = perc_destroy(h)
end program bond_perc_b200
