"""dl-poly_b200 -- B200-native short-range two-body path for DL_POLY (ccp5UK/dl-poly 5.1.0).

Scope: link-cell Verlet neighbour list (neighbours.F90::link_cell_pairs), the pair loops of
two_body.F90::two_body_forces (vdw.F90 tabulated/direct, ewald_spole.F90 real-space + exclusion correction), the
padding-driven rebuild test (vnl_check) and DL_POLY's domain decomposition halo/migration exchange.  Everything
numerical runs in hand-written CUDA (sm_100a) inside ``libdlpgpu.so`` (C ABI: include/dlpgpu.h); this package is the
host-side mirror of the reference call sites.  There is no CPU fallback: importing :mod:`.lib` raises if the CUDA
library is missing.

The directory name carries a hyphen (as the task layout prescribes), so import it through ``_pkg.load()`` at the
repo root (registers it as the module ``dl_poly_b200``).
"""
__version__ = "0.1.0"
