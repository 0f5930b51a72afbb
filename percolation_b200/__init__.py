"""percolation_b200 -- B200 (sm_100a) replacement for the percolation-realization hot path of
IsaiahSteinke/Percolation: occupancy generator, cluster labeling, spanning / cluster sizes,
Kirchhoff conductance.  The product is libperc_b200.so (hand-written CUDA behind the C-ABI of
include/perc_abi.h); this package only loads it.  No CPU fallback."""
from .lib import (BOND, MIXED, SITE, SQUARE, TRIANGULAR, Lattice, PercError, SO_PATH, SYMBOLS,  # noqa: F401
                  geom_bondlist, geom_nb, geom_nearestn, load)
