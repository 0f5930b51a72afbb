!===============================================================================
! sb_perc_b200.f90 -- the reference's PROGRAMs sb_perc and bs_perc (Fortran/Square/sb_perc.f, bs_perc.f):
! the mixed site/bond percolation line.  sb_perc: for ps = 0.59 + 0.01 k (42 points, :66-71), `iter` = 100
! trials each, occupy int(ps*t) sites and add bonds until the first spanning cluster; output row
! "sseed, bseed, ps, pb" with pb = 0 when it never spans (:363-369).  Seeds: srand(8811064), pseed(i) =
! int(rand(0)*10000000)+1; per ps point srand(pseed(ii)) and sseed(jj), bseed(jj) drawn INTERLEAVED
! (:94-112).  bs_perc is the mirror image (bonds fixed at pb = 0.30 + 0.01 k, sites added; set
! `sites_added = .true.`).  The add-until-spanning loops (Sq/sb_perc.f:232-357, Sq/bs_perc.f:237-380)
! are perc_first_span with the other element type held at its fill count.
! Not compile-tested in this image (no Fortran compiler).
!===============================================================================
program sb_perc_b200
  use iso_c_binding
  use perc_iface
  implicit none
  logical, parameter :: sites_added = .false.            ! .false.: sb_perc (bonds added)  .true.: bs_perc (sites added)
  integer(c_int32_t) :: m, n, t, pbc, lattice, nb, rc, seed, npts, iter, i, ii, jj, j, t1, t2
  integer(c_int32_t) :: kstar, maxcs, perccls, kfixed, minus1
  integer(c_int32_t), allocatable :: b(:), border(:), sorder(:), pseed(:), sseed(:), bseed(:)
  integer(c_int64_t) :: h
  double precision :: pfix
  real(c_float) :: f
  real :: rand

  if (sites_added) then
     open(unit=10, file='bs_perc.txt')
  else
     open(unit=10, file='sb_perc.txt')
  end if
  m = 50; n = 50; t = m*n; pbc = 0; lattice = PERC_SQUARE
  seed = 8811064                                          ! Sq/sb_perc.f:52
  if (sites_added) then
     npts = 71; iter = 1000                               ! Sq/bs_perc.f:66-75
  else
     npts = 42; iter = 100                                ! Sq/sb_perc.f:66-75
  end if
  minus1 = -1
  rc = perc_geom_nb(lattice, m, n, pbc, nb)
  allocate(b(2*nb), border(2*nb), sorder(t), pseed(100), sseed(1000), bseed(1000))
  rc = perc_geom_bondlist(lattice, m, n, pbc, b)
  call srand(seed)
  do i = 1, 100
     pseed(i) = int(rand(0)*10000000)+1                   ! :94-97
  end do
  rc = perc_create(h, lattice, m, n, pbc, 0)
  if (rc /= 0) stop 'perc_create failed (no CUDA device? there is no CPU fallback)'

  do ii = 1, npts
     if (sites_added) then
        pfix = 0.30d+00 + (0.01d+00*(ii-1))
     else
        pfix = 0.59d+00 + (0.01d+00*(ii-1))
     end if
     call srand(pseed(ii))                                ! :108-112
     do jj = 1, 1000
        sseed(jj) = int(rand(0)*10000000)+1
        bseed(jj) = int(rand(0)*10000000)+1
     end do
     do jj = 1, iter
        call srand(sseed(jj))                             ! site order, :151-162
        do i = 1, t
           sorder(i) = i
        end do
        do i = 1, t
           j = i + (t-i+1)*rand(0)
           t1 = sorder(i); sorder(i) = sorder(j); sorder(j) = t1
        end do
        border = b
        call srand(bseed(jj))                             ! bond order, :193-204
        do i = 1, nb
           j = i + (nb-i+1)*rand(0)
           t1 = border(i);    border(i) = border(j);       border(j) = t1
           t2 = border(nb+i); border(nb+i) = border(nb+j); border(nb+j) = t2
        end do
        rc = perc_set_site_order(h, sorder)
        rc = perc_set_bond_order(h, border)
        if (sites_added) then
           kfixed = pfix*nb
           rc = perc_set_fill(h, minus1, kfixed)
           rc = perc_first_span(h, PERC_MIXED, PERC_SITE, kstar, f, maxcs, perccls)
           write(10,111) sseed(jj), bseed(jj), dble(f), pfix          ! ps = 0 when it never spans (kstar = 0)
        else
           kfixed = pfix*t
           rc = perc_set_fill(h, kfixed, minus1)
           rc = perc_first_span(h, PERC_MIXED, PERC_BOND, kstar, f, maxcs, perccls)
           write(10,111) sseed(jj), bseed(jj), pfix, dble(f)          ! pb = 0 when it never spans
        end if
     end do
  end do
111 format(i10,",",i10,",",f12.9,",",f12.9)
  rc = perc_destroy(h)
end program sb_perc_b200
