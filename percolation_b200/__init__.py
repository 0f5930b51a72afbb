"""percolation_b200 -- B200 (sm_100a) replacement for the percolation-realization hot path of
IsaiahSteinke/Percolation: occupancy generator, cluster labeling, spanning / cluster sizes,
Kirchhoff conductance.  The product is libperc_b200.so (hand-written CUDA behind the C-ABI of
include/perc_abi.h); this package only loads it.  No CPU fallback."""
from .lib import (BOND, E_ARG, E_NOSPAN, E_ODD_M, E_SIZE, E_STATE, MIXED, SITE, SQUARE, TRIANGULAR,  # noqa: F401
                  Lattice, PercError, SlabLattice, SO_PATH, SYMBOLS, comm_unique_id, stitch_host, IFACE_WORDS,
                  geom_bondlist, geom_nb, geom_nearestn, load, matlab_variable_conductances)
