!===============================================================================
! bondc_b200.f90 -- the reference's PROGRAMs bond and bondc (Fortran/Square/bond.f, bondc.f and the
! Triangular twins): one bond-percolation realization at fraction pb, clusters = sets of bonds, and
! (bondc) the Kirchhoff conductance of the spanning cluster.  Same parameter block (Sq/bondc.f:67-92),
! seed 626504 (bond.f: 184489), shuffle of the bond list (Sq/bond.f:137-150), output bond.txt rows
! "b1, b2, label, j, c(j)" (Sq/bond.f:443-447) and the "Conductance: Gtop Gbot" line (Sq/bondc.f:593).
! Replaced: the O(nb^2) fill (Sq/bond.f:165-369), the spanning scan (:389-432) -> perc_bond; the dense
! G assembly + sprsin + linbcg + read-out (Sq/bondc.f:465-595) -> perc_conduct.
! Not compile-tested in this image (no Fortran compiler).
!===============================================================================
program bondc_b200
  use iso_c_binding
  use perc_iface
  implicit none
  integer(c_int32_t) :: m, n, t, pbc, lattice, nb, rc, seed, tbonds, i, j, t1, t2, iter
  integer(c_int32_t) :: maxcs, perccln, perccls
  integer(c_int32_t), allocatable :: b(:), border(:), b3(:), c(:)
  integer(c_int64_t) :: h
  double precision :: pb, Va, g0, Gtop, Gbot, err
  real :: rand

  open(unit=10, file='bond.txt')
  m = 50; n = 50; t = m*n; pbc = 0; pb = 0.50d+00        ! Sq/bondc.f:67-77
  Va = 1.0d+00; g0 = 1.0d+00                             ! :85-86
  lattice = PERC_SQUARE
  rc = perc_geom_nb(lattice, m, n, pbc, nb)
  allocate(b(2*nb), border(2*nb), b3(nb), c(t))
  rc = perc_geom_bondlist(lattice, m, n, pbc, b)          ! b(nb,2), Sq/bond.f:112-129
  border = b
  seed = 626504                                          ! Sq/bondc.f:81
  call srand(seed)
  do i = 1, nb                                           ! Sq/bond.f:142-150: both columns swap
     j = i + (nb-i+1)*rand(0)
     t1 = border(i);    border(i) = border(j);       border(j) = t1
     t2 = border(nb+i); border(nb+i) = border(nb+j); border(nb+j) = t2
  end do

  tbonds = pb*nb                                         ! Sq/bond.f:167
  rc = perc_create(h, lattice, m, n, pbc, 0)
  if (rc /= 0) stop 'perc_create failed (no CUDA device? there is no CPU fallback)'
  rc = perc_bond(h, border, tbonds, b3, c, maxcs, perccln, perccls)
  if (rc /= 0) stop 'perc_bond failed'
  if (perccln /= 0) then
     ! reference solver settings Sq/bondc.f:545 (tol 1e-8, itmax 2500), leak 1e-12 (:487), threshold 1e-10 (:576)
     rc = perc_conduct(h, perccln, Va, g0, 1.0d-12, 1.0d-8, 2500, 1.0d-10, Gtop, Gbot, iter, err)
     write(6,*) "Conductance:", Gtop, Gbot
  end if
  do j = 1, nb                                           ! Sq/bond.f:443-447; c() is indexed by the canonical label
     if (j <= t) then
        write(10,111) b(j), b(nb+j), b3(j), j, c(j)
     else
        write(10,111) b(j), b(nb+j), b3(j), j, 0
     end if
  end do
111 format(i10,",",i10,",",i10,",",i10,",",i10)
  rc = perc_destroy(h)
end program bondc_b200
