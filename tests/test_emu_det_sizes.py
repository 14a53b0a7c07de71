"""Determinant sizes between the configurations of BASELINE.json: every kernel selection of ``jq_launch_logdet`` gets a
network-level parity case (FermiNet on a single nucleus, full determinants of n = n_up + n_dn electrons):
thread per determinant (n <= 4), half-warp per determinant with synchronous slab loads (odd n or odd determinant
count) and with cp.async-staged slabs (n and the determinant count even), block per determinant group (n > 16).  The
reference has no such size sweep; the float64 oracle is the checker (sign bit-exact, scaled float32 tolerances)."""

import pytest

import helpers as H
import test_emu_edge_spins as E

# (nspins, nuclear charge, determinants)
CASES = [
    ((3, 2), 5.0, 3),    # n = 5, odd: synchronous loads
    ((3, 3), 6.0, 4),    # n = 6, 4 determinants: staged
    ((4, 3), 7.0, 4),    # n = 7, odd n with an even determinant count: synchronous loads
    ((5, 5), 10.0, 3),   # n = 10 with an odd determinant count: synchronous loads
    ((6, 6), 12.0, 2),   # n = 12: staged, one warp busy
    ((8, 8), 16.0, 4),   # n = 16: staged, the largest half-warp size
    ((8, 8), 16.0, 16),  # n = 16 with 16 determinants: the staging ring would not fit beside two blocks per SM -> synchronous loads
    ((9, 8), 17.0, 2),   # n = 17: block kernel
]
IDS = ["n%d_d%d" % (c[0][0] + c[0][1], c[2]) for c in CASES]


@pytest.mark.parametrize("nspins,charge,ndets", CASES, ids=IDS)
def test_determinant_sizes(nspins, charge, ndets):
    E.check(H.emu_runtime(), nspins, charge, "ferminet", ndets=ndets)
