"""JoinT ingestion, CPU side: the oracle's RING <-> NEST conversions against the HEALPix primer's
tables and round trips, and its he_udgrade restatement against the compiled reference's own he_udgrade
(src/healpix_extra.c:318-385, oracle/_ref/libgethi_ref.so)."""
import ctypes as C

import numpy as np
import pytest

from oracle.binding import Reference

# nest2ring for nside = 2 (HEALPix primer, fig. 4: NESTED index -> RING index)
NEST2RING_NSIDE2 = [13, 5, 4, 0, 15, 7, 6, 1, 17, 9, 8, 2, 19, 11, 10, 3, 28, 20, 27, 12, 30, 22, 21, 14, 32, 24, 23, 16, 34, 26, 25, 18,
                    44, 37, 36, 29, 45, 39, 38, 31, 46, 41, 40, 33, 47, 43, 42, 35]


def test_nest_ring_known_answers_and_round_trips(oracle):
    assert list(oracle.nest2ring(1, np.arange(12))) == list(range(12))
    assert list(oracle.nest2ring(2, np.arange(48))) == NEST2RING_NSIDE2
    for nside in (1, 2, 4, 8, 32):
        npix = 12 * nside * nside
        ring = oracle.nest2ring(nside, np.arange(npix))
        assert sorted(ring) == list(range(npix))                       # a bijection
        assert np.array_equal(oracle.ring2nest(nside, ring), np.arange(npix))
    rng = np.random.default_rng(0)
    for nside in (256, 1024, 8192):
        pix = rng.integers(0, 12 * nside * nside, 3000)
        assert np.array_equal(oracle.ring2nest(nside, oracle.nest2ring(nside, pix)), pix)


def test_nest_children_are_neighbours_on_the_sphere(oracle):
    """The four children of a NEST pixel lie inside their parent: their centres are closer to the parent's centre
    than a parent pixel's size (ties the bit interleave to the geometry, via the oracle's pix2vec_ring)."""
    nside = 16
    for parent in (0, 5, 100, 12 * nside * nside - 1, 777):
        pr = int(oracle.nest2ring(nside, [parent])[0])
        vp = oracle.pix2vec_ring(nside, pr)
        for ch in range(4):
            cr = int(oracle.nest2ring(2 * nside, [4 * parent + ch])[0])
            vc = oracle.pix2vec_ring(2 * nside, cr)
            assert np.arccos(np.clip(np.dot(vp, vc), -1, 1)) < 1.2 * np.sqrt(4 * np.pi / (12 * nside * nside))


def test_udgrade_properties(oracle):
    rng = np.random.default_rng(1)
    m = rng.normal(size=12 * 32 * 32).astype(np.float32)
    assert np.array_equal(oracle.udgrade(m, 32), m)
    up = oracle.udgrade(m, 128)
    assert np.array_equal(oracle.udgrade(up, 32), m)                   # replicate, then average identical values
    down = oracle.udgrade(m, 8)
    assert abs(float(down.astype(np.float64).mean()) - float(m.astype(np.float64).mean())) < 1e-7
    # RING and NEST orderings agree through the index maps
    nest = m[oracle.nest2ring(32, np.arange(m.size))]
    down_nest = oracle.udgrade(nest, 8, nest=True)
    assert np.array_equal(down_nest, down[oracle.nest2ring(8, np.arange(down.size))])


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("nside_in,nside_out,nest", [(64, 16, 0), (16, 64, 0), (32, 32, 0), (64, 8, 1), (8, 32, 1), (128, 64, 0)])
def test_udgrade_restatement_equals_the_reference(oracle, nside_in, nside_out, nest):
    ref = Reference()
    rng = np.random.default_rng(nside_in * 7 + nside_out)
    m = (rng.normal(size=12 * nside_in * nside_in) * 10 ** rng.uniform(-3, 3)).astype(np.float32)
    out = np.zeros(12 * nside_out * nside_out, np.float32)
    ref.lib.he_udgrade.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_int]
    ref.lib.he_udgrade.restype = None
    ref.lib.he_udgrade(m.ctypes.data, nside_in, out.ctypes.data, nside_out, nest)
    assert np.array_equal(out, oracle.udgrade(m, nside_out, nest=bool(nest)))
