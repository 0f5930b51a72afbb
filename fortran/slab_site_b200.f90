!===============================================================================
! slab_site_b200.f90 -- one square-lattice site realization at L = 65536 on 8 GPUs
! (BASELINE configs[4]); no counterpart in the reference, whose arrays stop at
! t = 10^6 (Fortran/Square/site.f:37).  One MPI rank per GPU; the lattice is cut
! into row slabs, the library labels each slab, stitches the clusters over NCCL
! and solves the Kirchhoff problem with halo exchange + all-reduced dot products.
! Not compile-tested here (no Fortran compiler / MPI in the image).
! Build: mpifort -O2 slab_site_b200.f90 perc_iface.o -L.. -lperc_b200
!===============================================================================
program slab_site_b200
  use iso_c_binding
  use perc_iface
  use mpi
  implicit none
  integer(c_int32_t), parameter :: m = 65536, n = 65536, pbc = 0
  integer(c_int64_t) :: h, t, ks, ncl, maxcs, maxcn, nspan, ids(16), sizes(16)
  integer(c_int32_t) :: rc, rank, nranks, ierr, ya, yb, ns, iter
  integer(c_int8_t) :: id(128)
  real(c_double) :: ps, Gtop, Gbot, err

  call MPI_Init(ierr)
  call MPI_Comm_rank(MPI_COMM_WORLD, rank, ierr)
  call MPI_Comm_size(MPI_COMM_WORLD, nranks, ierr)
  if (rank == 0) rc = perc_comm_unique_id(id)
  call MPI_Bcast(id, 128, MPI_BYTE, 0, MPI_COMM_WORLD, ierr)

  rc = perc_create_slab(h, PERC_SQUARE, m, n, pbc, rank, nranks, rank)     ! device = local rank
  if (rc /= 0) stop 'perc_create_slab'
  rc = perc_comm_init(h, id)
  rc = perc_slab_rows(h, ya, yb)

  ps = 0.60d0                                   ! site threshold 0.593 (Sq/site.f:24-25)
  t = int(m, c_int64_t) * n
  ks = int(ps * dble(t), c_int64_t)             ! tsites = ps*t, truncated (Sq/site.f:164)
  rc = perc_generate_i8(h, 20240611_c_int64_t, 0_c_int64_t, ks, -1_c_int64_t)
  rc = perc_label(h, PERC_SITE)
  rc = perc_summary_i8(h, ncl, maxcs, maxcn, nspan)
  rc = perc_span_i8(h, 16, ns, ids, sizes)
  if (rank == 0) write (6, *) 'clusters', ncl, ' largest', maxcs, ' spanning', nspan
  if (ns > 0) then
    rc = perc_conduct_g(h, 0, 1.0d0, 1.0d0, 1.0d-12, 1.0d-8, 200000, 1.0d-10, Gtop, Gbot, iter, err)
    if (rank == 0) write (6, '(a,2f14.9,i8,es10.2)') ' Gtop, Gbot, iter, err ', Gtop, Gbot, iter, err
  end if
  rc = perc_destroy(h)
  call MPI_Finalize(ierr)
end program
