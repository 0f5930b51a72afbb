!===============================================================================
! perc_iface.f90 -- ISO_C_BINDING interface to libperc_b200.so (include/perc_abi.h)
!
! The reference drivers (Fortran/Square/*.f, Fortran/Triangular/*.f) keep their
! structure: parameter block, seed derivation, shuffle (srand/rand), output
! files.  Their hot code blocks are replaced by the calls declared here:
!
!   shuffle + "occupy order(1..k)"   Sq/site.f:131-176    -> perc_set_site_order / perc_set_fill
!                                    Sq/bond.f:137-167    -> perc_set_bond_order
!   (or, without the sequential RNG) Fortran/permute.f    -> perc_generate (Philox, exact count)
!   main labeling loop               Sq/site.f:162-289    -> perc_site / perc_label
!                                    Sq/bond.f:165-369    -> perc_bond
!                                    Sq/sitebond.f:187-400-> perc_sitebond
!   spanning scan                    Sq/site.f:309-344    -> perccln/perccls of the above, perc_span
!   fill-until-spanning loops        Sq/site_perc.f:133-254 -> perc_first_span
!   conductance block                Sq/bondc.f:465-595   -> perc_conduct
!
! All scalars by reference (no VALUE), INTEGER = c_int32_t, DOUBLE PRECISION =
! c_double, arrays 1-based / column-major exactly as the reference declares them.
! NOTE: this image has no Fortran compiler (gfortran/flang/nvfortran absent), so
! this module is not compile-tested here; tests/abi_c_driver.c exercises the
! identical by-reference calling convention from C.
! Build:  gfortran -O2 -c perc_iface.f90 ; gfortran -O2 site_b200.f90 perc_iface.o -L.. -lperc_b200
!===============================================================================
module perc_iface
  use iso_c_binding
  implicit none

  integer(c_int32_t), parameter :: PERC_SQUARE = 1, PERC_TRIANGULAR = 2
  integer(c_int32_t), parameter :: PERC_SITE = 1, PERC_BOND = 2, PERC_MIXED = 3

  interface
    integer(c_int32_t) function perc_geom_nb(lattice, m, n, pbc, nb) bind(C, name="perc_geom_nb")
      import :: c_int32_t
      integer(c_int32_t), intent(in) :: lattice, m, n, pbc
      integer(c_int32_t), intent(out) :: nb
    end function

    integer(c_int32_t) function perc_geom_bondlist(lattice, m, n, pbc, b) bind(C, name="perc_geom_bondlist")
      import :: c_int32_t
      integer(c_int32_t), intent(in) :: lattice, m, n, pbc
      integer(c_int32_t), intent(out) :: b(*)            ! b(nb,2)
    end function

    integer(c_int32_t) function perc_geom_nearestn(lattice, m, n, pbc, rn, nn) bind(C, name="perc_geom_nearestn")
      import :: c_int32_t
      integer(c_int32_t), intent(in) :: lattice, m, n, pbc, rn
      integer(c_int32_t), intent(out) :: nn(6)
    end function

    integer(c_int32_t) function perc_create(h, lattice, m, n, pbc, device) bind(C, name="perc_create")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(out) :: h
      integer(c_int32_t), intent(in) :: lattice, m, n, pbc, device
    end function

    integer(c_int32_t) function perc_destroy(h) bind(C, name="perc_destroy")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
    end function

    integer(c_int32_t) function perc_sync(h) bind(C, name="perc_sync")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
    end function

    integer(c_int32_t) function perc_set_site_order(h, order) bind(C, name="perc_set_site_order")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: order(*)          ! order(t)
    end function

    integer(c_int32_t) function perc_set_bond_order(h, border) bind(C, name="perc_set_bond_order")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: border(*)         ! border(nb,2)
    end function

    integer(c_int32_t) function perc_set_fill(h, ks, kb) bind(C, name="perc_set_fill")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: ks, kb            ! -1 keeps the current value
    end function

    integer(c_int32_t) function perc_set_occupancy(h, socc, bocc) bind(C, name="perc_set_occupancy")
      import :: c_int32_t, c_int64_t, c_ptr
      integer(c_int64_t), intent(in) :: h
      type(c_ptr), value :: socc, bocc                    ! c_loc(byte array) or c_null_ptr
    end function

    integer(c_int32_t) function perc_generate(h, seed, stream, ks, kb) bind(C, name="perc_generate")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h, seed, stream
      integer(c_int32_t), intent(in) :: ks, kb
    end function

    integer(c_int32_t) function perc_get_occupancy(h, socc, bocc) bind(C, name="perc_get_occupancy")
      import :: c_int32_t, c_int64_t, c_ptr
      integer(c_int64_t), intent(in) :: h
      type(c_ptr), value :: socc, bocc
    end function

    integer(c_int32_t) function perc_label(h, kind) bind(C, name="perc_label")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: kind
    end function

    integer(c_int32_t) function perc_summary(h, ncl, maxcs, maxcn, nspan) bind(C, name="perc_summary")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int64_t), intent(out) :: ncl
      integer(c_int32_t), intent(out) :: maxcs, maxcn, nspan
    end function

    integer(c_int32_t) function perc_get_site_labels(h, s) bind(C, name="perc_get_site_labels")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(out) :: s(*)
    end function

    integer(c_int32_t) function perc_get_bond_labels(h, b3) bind(C, name="perc_get_bond_labels")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(out) :: b3(*)
    end function

    integer(c_int32_t) function perc_get_sizes(h, c) bind(C, name="perc_get_sizes")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(out) :: c(*)
    end function

    integer(c_int32_t) function perc_span(h, max_ids, nspan, ids, sizes) bind(C, name="perc_span")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: max_ids
      integer(c_int32_t), intent(out) :: nspan, ids(*), sizes(*)
    end function

    integer(c_int32_t) function perc_hist(h, nbins, hist) bind(C, name="perc_hist")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: nbins
      integer(c_int64_t), intent(out) :: hist(*)
    end function

    integer(c_int32_t) function perc_hist_log2(h, nbins, hist) bind(C, name="perc_hist_log2")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: nbins
      integer(c_int64_t), intent(out) :: hist(*)
    end function

    integer(c_int32_t) function perc_site(h, order, k, s, c, maxcs, perccln, perccls) bind(C, name="perc_site")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: order(*), k
      integer(c_int32_t), intent(out) :: s(*), c(*), maxcs, perccln, perccls
    end function

    integer(c_int32_t) function perc_bond(h, border, k, b3, c, maxcs, perccln, perccls) bind(C, name="perc_bond")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: border(*), k
      integer(c_int32_t), intent(out) :: b3(*), c(*), maxcs, perccln, perccls
    end function

    integer(c_int32_t) function perc_sitebond(h, sorder, ks, border, kb, s, b3, c, maxcs, perccln, perccls) &
        bind(C, name="perc_sitebond")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: sorder(*), ks, border(*), kb
      integer(c_int32_t), intent(out) :: s(*), b3(*), c(*), maxcs, perccln, perccls
    end function

    integer(c_int32_t) function perc_first_span(h, kind, which, kstar, f, maxcs, perccls) bind(C, name="perc_first_span")
      import :: c_int32_t, c_int64_t, c_float
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: kind, which
      integer(c_int32_t), intent(out) :: kstar, maxcs, perccls
      real(c_float), intent(out) :: f
    end function

    integer(c_int32_t) function perc_conduct(h, cluster_id, Va, g0, gleak, tol, itmax, read_thresh, &
                                             Gtop, Gbot, iter, err) bind(C, name="perc_conduct")
      import :: c_int32_t, c_int64_t, c_double
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: cluster_id, itmax
      real(c_double), intent(in) :: Va, g0, gleak, tol, read_thresh
      real(c_double), intent(out) :: Gtop, Gbot, err
      integer(c_int32_t), intent(out) :: iter
    end function

    ! perc_conduct starting from the voltages of the previous sweep point (linbcg's x is in/out, Sq/bondc.f:759-763)
    integer(c_int32_t) function perc_conduct_warm(h, cluster_id, Va, g0, gleak, tol, itmax, read_thresh, &
                                                  Gtop, Gbot, iter, err) bind(C, name="perc_conduct_warm")
      import :: c_int32_t, c_int64_t, c_double
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: cluster_id, itmax
      real(c_double), intent(in) :: Va, g0, gleak, tol, read_thresh
      real(c_double), intent(out) :: Gtop, Gbot, err
      integer(c_int32_t), intent(out) :: iter
    end function

    ! the same solve for the p-sweep drivers (Sq/bond_cond.f:392-485 write only pb, Gbot, Gtop, avg):
    ! identical Gtop / Gbot / iter / err, interior voltages not formed
    integer(c_int32_t) function perc_conduct_g(h, cluster_id, Va, g0, gleak, tol, itmax, read_thresh, &
                                               Gtop, Gbot, iter, err) bind(C, name="perc_conduct_g")
      import :: c_int32_t, c_int64_t, c_double
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: cluster_id, itmax
      real(c_double), intent(in) :: Va, g0, gleak, tol, read_thresh
      real(c_double), intent(out) :: Gtop, Gbot, err
      integer(c_int32_t), intent(out) :: iter
    end function

    ! ---- trial loops on the device (do ii = 1, numtrials: Sq/site_perc.f:87, Sq/bond_cond.f:123) ----
    integer(c_int32_t) function perc_batch(h, kind, nreal, seed, stream0, ks, kb, nbins, hist, stats) &
        bind(C, name="perc_batch")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h, seed, stream0
      integer(c_int32_t), intent(in) :: kind, nreal, ks, kb, nbins
      integer(c_int64_t), intent(out) :: hist(*), stats(16)
    end function

    integer(c_int32_t) function perc_batch_conduct(h, kind, nreal, seed, stream0, ks, kb, Va, g0, gleak, tol, itmax, &
                                                   read_thresh, G, iters, stats) bind(C, name="perc_batch_conduct")
      import :: c_int32_t, c_int64_t, c_double
      integer(c_int64_t), intent(in) :: h, seed, stream0
      integer(c_int32_t), intent(in) :: kind, nreal, ks, kb, itmax
      real(c_double), intent(in) :: Va, g0, gleak, tol, read_thresh
      real(c_double), intent(out) :: G(2, *)             ! Gtop, Gbot per realization
      integer(c_int32_t), intent(out) :: iters(*)
      integer(c_int64_t), intent(out) :: stats(16)
    end function

    ! ---- several GPUs: one process (MPI rank) per GPU ------------------------------------------------
    ! NCCL bootstrap: rank 0 calls perc_comm_unique_id, MPI_Bcast the 128 bytes, every rank initialises
    integer(c_int32_t) function perc_comm_unique_id(id128) bind(C, name="perc_comm_unique_id")
      import :: c_int32_t, c_int8_t
      integer(c_int8_t), intent(out) :: id128(128)
    end function

    ! mode 1: independent realizations per rank, one reduction of the statistics at the end
    integer(c_int32_t) function perc_comm_init_rank(h, nranks, rank, id128) bind(C, name="perc_comm_init_rank")
      import :: c_int32_t, c_int64_t, c_int8_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: nranks, rank
      integer(c_int8_t), intent(in) :: id128(128)
    end function

    integer(c_int32_t) function perc_allreduce_stats(h, ni, ivals, nd, dvals) bind(C, name="perc_allreduce_stats")
      import :: c_int32_t, c_int64_t, c_double
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: ni, nd
      integer(c_int64_t), intent(inout) :: ivals(*)
      real(c_double), intent(inout) :: dvals(*)
    end function

    ! mode 2: ONE lattice decomposed into row slabs (rank r holds rows n*r/nranks .. n*(r+1)/nranks - 1)
    integer(c_int32_t) function perc_create_slab(h, lattice, m, n, pbc, device, nranks, rank) bind(C, name="perc_create_slab")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(out) :: h
      integer(c_int32_t), intent(in) :: lattice, m, n, pbc, device, nranks, rank
    end function

    integer(c_int32_t) function perc_comm_init(h, id128) bind(C, name="perc_comm_init")
      import :: c_int32_t, c_int64_t, c_int8_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int8_t), intent(in) :: id128(128)
    end function

    integer(c_int32_t) function perc_slab_rows(h, ya, yb) bind(C, name="perc_slab_rows")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(out) :: ya, yb
    end function

    integer(c_int32_t) function perc_generate_i8(h, seed, stream, ks, kb) bind(C, name="perc_generate_i8")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h, seed, stream, ks, kb
    end function

    integer(c_int32_t) function perc_summary_i8(h, ncl, maxcs, maxcn, nspan) bind(C, name="perc_summary_i8")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int64_t), intent(out) :: ncl, maxcs, maxcn, nspan
    end function

    integer(c_int32_t) function perc_span_i8(h, max_ids, nspan, ids, sizes) bind(C, name="perc_span_i8")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: max_ids
      integer(c_int32_t), intent(out) :: nspan
      integer(c_int64_t), intent(out) :: ids(*), sizes(*)
    end function

    integer(c_int32_t) function perc_get_site_labels_i8(h, s) bind(C, name="perc_get_site_labels_i8")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int64_t), intent(out) :: s(*)            ! s((yb-ya)*m): lattice-wide labels of the owned rows
    end function

    integer(c_int32_t) function perc_get_voltage(h, Vint) bind(C, name="perc_get_voltage")
      import :: c_int32_t, c_int64_t, c_double
      integer(c_int64_t), intent(in) :: h
      real(c_double), intent(out) :: Vint(*)
    end function

    ! iteration kernels of the conductance solve: mode 0 (default) = the one-pass kernel wherever it applies
    ! (perc_conduct_g, one GPU, pbc = 0), mode 1 = always the two-kernel form (iter is linbcg's count exactly)
    integer(c_int32_t) function perc_set_solver(h, mode) bind(C, name="perc_set_solver")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: mode
    end function

    ! re-labeling along a sweep: after perc_set_fill raised the fill of a labeled handle only the added elements are united;
    ! incremental = 1 if that pass ran, 0 if the full perc_label pass was needed
    integer(c_int32_t) function perc_label_incremental(h, kind, incremental) bind(C, name="perc_label_incremental")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: kind
      integer(c_int32_t), intent(out) :: incremental
    end function

    ! the reference programs' text files from the current labeling, in their own formats (which = 1 site.txt, 2 bond.txt,
    ! 3 sbsite.txt, 4 sbbond.txt, 5 bondlist.txt):  ierr = perc_write_txt(h, 1, 'site.txt', len('site.txt'))
    integer(c_int32_t) function perc_write_txt(h, which, path, pathlen) bind(C, name="perc_write_txt")
      import :: c_int32_t, c_int64_t, c_char
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: which, pathlen
      character(kind=c_char), intent(in) :: path(*)
    end function

    ! per-bond conductances (MATLAB/ConductCalc.m condtype = 2: g0*rand on the conducting bonds): w(nb) in bond-list
    ! order, used for the bonds that conduct; perc_clear_bond_conductance drops them
    integer(c_int32_t) function perc_set_bond_conductance(h, w) bind(C, name="perc_set_bond_conductance")
      import :: c_int32_t, c_int64_t, c_double
      integer(c_int64_t), intent(in) :: h
      real(c_double), intent(in) :: w(*)
    end function
    integer(c_int32_t) function perc_clear_bond_conductance(h) bind(C, name="perc_clear_bond_conductance")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
    end function

    integer(c_int32_t) function perc_solver_used(h, fused) bind(C, name="perc_solver_used")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(out) :: fused
    end function

    integer(c_int32_t) function perc_launch_count(h, count) bind(C, name="perc_launch_count")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int64_t), intent(out) :: count
    end function

    integer(c_int32_t) function perc_phase_ms(h, nphase, ms) bind(C, name="perc_phase_ms")
      import :: c_int32_t, c_int64_t, c_float
      integer(c_int64_t), intent(in) :: h
      integer(c_int32_t), intent(in) :: nphase
      real(c_float), intent(out) :: ms(*)
    end function

    integer(c_int32_t) function perc_stream(h, stream) bind(C, name="perc_stream")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: h
      integer(c_int64_t), intent(out) :: stream
    end function
  end interface
end module perc_iface
