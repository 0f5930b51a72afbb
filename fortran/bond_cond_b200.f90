!===============================================================================
! bond_cond_b200.f90 -- the reference's PROGRAM bond_cond (Fortran/Square/bond_cond.f)
! on top of libperc_b200.  Kept from the reference: parameter block (:47-60), seed table
! tseed(i) = int(rand(0)*10000000)+1 (:64-70), sweep table pbarr(1)=0.49d0,
! pbarr(i)=pbarr(i-1)+5d-3, nbarr(i)=pbarr(i)*nb (:84-97), bond-list enumeration and shuffle
! (:140-180), output rows "pb,Gbot,Gtop,avg" with format f12.9 (:481-482,505).
! Replaced: the add-one-bond-and-rescan loop (:208-370) and the dense conductance block
! (:392-476).  Because the bond order is uploaded once as a rank table, each sweep point is
! perc_set_fill(kb = nbarr(jj)) + perc_label_incremental + perc_conduct -- the nesting property of the
! reference's single fill order is preserved (App. B: the stall on duplicate nbarr entries is
! not reproduced; jj advances past duplicates).
! Not compile-tested in this image (no Fortran compiler).
!===============================================================================
program bond_cond_b200
  use iso_c_binding
  use perc_iface
  implicit none
  integer(c_int32_t) :: m, n, t, pbc, lattice, device, rc, nb
  integer(c_int32_t) :: numtrials, seed, i, ii, j, jj, keep, itmax, iter, cid
  integer(c_int32_t) :: maxcs, perccln, perccls, kstar, nspan, inc
  integer(c_int32_t) :: btemp(2), nbarr(250), ids(16), sizes(16)
  integer(c_int32_t), allocatable :: b(:,:), border(:,:), b3(:), c(:), tseed(:)
  integer(c_int64_t) :: h, ncl
  real(c_double) :: Va, g0, gleak, tol, thr, Gtop, Gbot, err, pbarr(250)
  real(c_float) :: f
  real :: rand, pb, pc

  open(unit=10, file='bondcond.txt')
  m = 10; n = 10; t = m*n; pbc = 0             ! Sq/bond_cond.f:47-55
  Va = 1.00d+00; g0 = 1.00d+00                 ! :59-60
  gleak = 1.00d-12; tol = 1.00d-08; itmax = 2500; thr = 1.00d-10   ! :408,466,497
  lattice = PERC_SQUARE; device = 0; keep = -1
  numtrials = 1; seed = 58302                  ! :64-65
  allocate(tseed(1000))
  call srand(seed)
  do i = 1, 1000
     tseed(i) = int(rand(0)*10000000)+1        ! :68-70
  end do
  rc = perc_geom_nb(lattice, m, n, pbc, nb)
  allocate(b(nb,2), border(nb,2), b3(nb), c(t))
  pbarr = 0.00d+00; nbarr = 0
  pbarr(1) = 0.49d+00
  do i = 2, 103
     pbarr(i) = pbarr(i-1)+5.00d-03             ! :89-92
  end do
  do i = 1, 250
     nbarr(i) = pbarr(i)*nb                     ! :93-94 (fp64 product truncated)
  end do
  rc = perc_geom_bondlist(lattice, m, n, pbc, b)     ! replaces :140-160 (nearestn enumeration)
  rc = perc_create(h, lattice, m, n, pbc, device)
  if (rc /= 0) stop 'perc_create failed (no CUDA device? there is no CPU fallback)'

  do ii = 1, numtrials
     write(10,*) "Trial #", ii
     write(10,*) "Random number seed:", tseed(ii)
     border = b
     call srand(tseed(ii))
     do i = 1, nb                               ! :170-180
        j = i + (nb-i+1)*rand(0)
        btemp(1) = border(i,1); btemp(2) = border(i,2)
        border(i,1) = border(j,1); border(i,2) = border(j,2)
        border(j,1) = btemp(1); border(j,2) = btemp(2)
     end do
     rc = perc_set_bond_order(h, border)
     ! pc = fraction at which a spanning cluster first appears (:381)
     rc = perc_first_span(h, PERC_BOND, PERC_BOND, kstar, f, maxcs, perccls)
     pc = f
     jj = 1
     do while (jj <= 103)
        if (jj > 1) then
           if (nbarr(jj) == nbarr(jj-1)) then   ! App. B: do not stall on duplicates
              jj = jj + 1
              cycle
           end if
        end if
        rc = perc_set_fill(h, keep, nbarr(jj))
        rc = perc_label_incremental(h, PERC_BOND, inc)   ! only the bonds added since the last point are united
        rc = perc_span(h, 16, nspan, ids, sizes)
        pb = real(nbarr(jj))/real(nb)
        if (nspan > 0) then
           cid = ids(1)
           rc = perc_conduct(h, cid, Va, g0, gleak, tol, itmax, thr, Gtop, Gbot, iter, err)
        else
           Gbot = 0.00d+00; Gtop = 0.00d+00
        end if
        write(6,111) pb, Gbot, Gtop, ((Gbot+Gtop)/2)
        write(10,111) pb, Gbot, Gtop, ((Gbot+Gtop)/2)
        jj = jj + 1
     end do
     write(10,*) "pc =", pc
     write(10,*) "------------------------------"
  end do
111 format(f12.9,",",f12.9,",",f12.9,",",f12.9)
  rc = perc_destroy(h)
end program bond_cond_b200
