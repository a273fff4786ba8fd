"""JoinT ingestion on the device against the oracle: RING <-> NEST, he_udgrade and merge_maps are integer /
fixed-order float arithmetic, so the device must be bit-identical (src/main_jt.c:98-211, src/healpix_extra.c:318-385)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gh(tables_nu64):
    from crime_b200 import GetHI, params_from_tables
    p = params_from_tables(tables_nu64, n_grid=64, n_side=64, seed=11)
    with GetHI(p) as g:
        yield g


@pytest.mark.parametrize("nside", [1, 2, 16, 256, 2048, 8192])
def test_nest_ring_on_device_equals_oracle(gh, oracle, nside):
    rng = np.random.default_rng(nside)
    npix = 12 * nside * nside
    pix = np.arange(npix) if npix <= 5000 else np.unique(np.concatenate([rng.integers(0, npix, 4000), [0, npix - 1, npix // 2]]))
    ring = gh.nest_ring(nside, pix, to_ring=True)
    assert np.array_equal(ring, oracle.nest2ring(nside, pix))
    assert np.array_equal(gh.nest_ring(nside, ring, to_ring=False), pix)
    assert np.array_equal(gh.nest_ring(nside, pix, to_ring=False), oracle.ring2nest(nside, pix))


@pytest.mark.parametrize("nside_in,nside_out,nest", [(64, 16, False), (16, 64, False), (32, 32, False), (128, 8, False), (64, 16, True),
                                                     (8, 32, True)])
def test_udgrade_on_device_is_bit_identical(gh, oracle, nside_in, nside_out, nest):
    rng = np.random.default_rng(nside_in + 3 * nside_out)
    stack = (rng.normal(size=(3, 12 * nside_in * nside_in)) * 10 ** rng.uniform(-4, 4, (3, 1))).astype(np.float32)
    out = gh.udgrade(stack, nside_out, nest=nest)
    for k in range(3):
        assert np.array_equal(out[k], oracle.udgrade(stack[k], nside_out, nest=nest))
    with pytest.raises(Exception):
        gh.udgrade(stack[0][: 12 * 9], 3)                                  # nside must be a power of two


def test_merge_maps_takes_the_cosmological_signal_from_the_device(gh, oracle):
    """merge_maps with the signal left on the device by run(), two foreground stacks from the host and the
    polarisation-leakage factor on one of them: equals the reference's order of float operations exactly."""
    cosmo = np.array(gh.run(), copy=True)                                  # [n_shells][npix], also still on the device
    n_nu, npix = cosmo.shape
    rng = np.random.default_rng(5)
    fg1 = (rng.lognormal(size=(n_nu, npix)) * 1e3).astype(np.float32)
    fg2 = rng.normal(size=(n_nu, npix)).astype(np.float32)
    leak = 0.013
    for nside_out in (16, 64, 128):
        out = gh.jt_merge_maps([None, fg1, fg2, fg2], nside_out, scale=[1.0, 1.0, leak, 1.0])
        for s in (0, n_nu // 2, n_nu - 1):
            acc = np.zeros(npix, np.float32)
            acc += cosmo[s]
            acc += fg1[s]
            acc += (fg2[s].astype(np.float64) * leak).astype(np.float32)   # map_read[ii] *= leakage (float *= double)
            acc += fg2[s]
            assert np.array_equal(out[s], oracle.udgrade(acc, nside_out))
    # the signal alone, same resolution: the maps themselves
    assert np.array_equal(gh.jt_merge_maps([None], 64), cosmo)
