"""The FITS maps this repository writes, read back by the REFERENCE's own reader: he_read_healpix_map
(/root/reference/src/healpix_extra.c:166-224, compiled unmodified into oracle/_ref/libgethi_ref.so) is what JoinT and
every downstream user of the reference open GetHI's output with -- HDU 2, NAXIS / NAXISn, NSIDE, ORDERING, column 1 as
npix floats, NEST -> RING reordering.  cfitsio is not installed; the reader behind the cfitsio names
(oracle/shim/shim_fitsio.c) is written against the FITS standard, independently of host/io.c's writer, and is itself
checked here on layouts neither writer of this repository produces (1024E rows, a double column, a further HDU).  CPU only."""
import ctypes as C

import numpy as np
import pytest

from crime_b200 import host
from oracle.binding import Reference

pytestmark = pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref():
    r = Reference()
    r.lib.he_read_healpix_map.argtypes = [C.c_char_p, C.POINTER(C.c_long), C.c_int]
    r.lib.he_read_healpix_map.restype = C.POINTER(C.c_float)
    r.lib.he_write_healpix_map.argtypes = [C.POINTER(C.POINTER(C.c_float)), C.c_int, C.c_long, C.c_char_p]
    r.lib.he_write_healpix_map.restype = None
    return r


def ref_read(ref, path, nfield=0):
    nside = C.c_long(-1)
    p = ref.lib.he_read_healpix_map(str(path).encode(), C.byref(nside), nfield)
    n = 12 * nside.value * nside.value
    return np.ctypeslib.as_array(p, shape=(n,)).copy(), nside.value


def card(key, val, string=False):
    v = f"'{val:<8s}'" if string else f"{val:>20}"
    return f"{key:<8s}= {v:<20s}".ljust(80).encode()


def fits_bytes(cols, nside, ordering, rep, extra_hdu=False):
    """A HEALPix FITS file assembled by hand: cols = list of (code, array); every row holds `rep` elements per column."""
    def block(cards):
        b = b"".join(cards) + "END".ljust(80).encode()
        return b + b" " * (-len(b) % 2880)
    out = block([card("SIMPLE", "T"), card("BITPIX", 8), card("NAXIS", 0), card("EXTEND", "T")])
    npix = 12 * nside * nside
    nrows = npix // rep
    width = sum(rep * (4 if c == "E" else 8) for c, _ in cols)
    cards = [card("XTENSION", "BINTABLE", True), card("BITPIX", 8), card("NAXIS", 2), card("NAXIS1", width), card("NAXIS2", nrows),
             card("PCOUNT", 0), card("GCOUNT", 1), card("TFIELDS", len(cols))]
    for i, (c, _) in enumerate(cols):
        cards += [card(f"TTYPE{i + 1}", "TQU"[i], True), card(f"TFORM{i + 1}", f"{rep}{c}", True)]
    cards += [card("PIXTYPE", "HEALPIX", True), card("ORDERING", ordering, True), card("NSIDE", nside)]
    rows = np.concatenate([np.asarray(a, ">f4" if c == "E" else ">f8").reshape(nrows, rep).view(np.uint8).reshape(nrows, -1) for c, a in cols], axis=1)
    data = rows.tobytes()
    out += block(cards) + data + b"\0" * (-len(data) % 2880)
    if extra_hdu:  # an image extension behind the table (the reference always opens HDU 2, healpix_extra.c:180)
        img = np.arange(7, dtype=">i2").tobytes()
        out += block([card("XTENSION", "IMAGE", True), card("BITPIX", 16), card("NAXIS", 1), card("NAXIS1", 7), card("PCOUNT", 0),
                      card("GCOUNT", 1)]) + img + b"\0" * (-len(img) % 2880)
    return out


@pytest.mark.parametrize("nside", [1, 8, 64])
def test_files_of_the_c_host_through_the_reference_reader(ref, tmp_path, nside):
    """host/io.c (write_maps' per-shell writer) -> the reference's he_read_healpix_map: same nside, same pixels, RING."""
    m = np.random.default_rng(nside).standard_normal(12 * nside * nside).astype(np.float32)
    f = tmp_path / "map_001.fits"
    assert host.write_healpix_map(f, m, nside) == 0
    back, ns = ref_read(ref, f)
    assert ns == nside and np.array_equal(back.view(np.uint32), m.view(np.uint32))


def test_reference_writer_and_c_host_writer_read_back_alike(ref, tmp_path):
    """The reference's own he_write_healpix_map (over the shim's writer) and host/io.c produce files the reference reads
    back identically, and our reader (crime_b200.host.read_healpix_map) agrees on both."""
    nside = 16
    m = np.random.default_rng(3).standard_normal(12 * nside * nside).astype(np.float32)
    col = m.ctypes.data_as(C.POINTER(C.c_float))
    cols = (C.POINTER(C.c_float) * 1)(col)
    ref.lib.he_write_healpix_map(cols, 1, nside, str(tmp_path / "ref.fits").encode())
    assert host.write_healpix_map(tmp_path / "own.fits", m, nside) == 0
    a, na = ref_read(ref, tmp_path / "ref.fits")
    b, nb = ref_read(ref, tmp_path / "own.fits")
    assert na == nb == nside and np.array_equal(a, m) and np.array_equal(b, m)
    for f in ("ref.fits", "own.fits"):
        back, hdr = host.read_healpix_map(tmp_path / f)
        assert np.array_equal(back, m) and hdr["ORDERING"] == "RING" and int(hdr["NSIDE"]) == nside


def test_shim_reader_on_layouts_this_repository_never_writes(ref, oracle, tmp_path):
    """healpy-style 1024-element rows, a second and third column, a double column, a further HDU behind the table, and
    NESTED ordering (which he_read_healpix_map turns into RING through nest2ring)."""
    nside = 32
    npix = 12 * nside * nside
    rng = np.random.default_rng(9)
    t, q, u = (rng.standard_normal(npix).astype(np.float32) for _ in range(3))
    (tmp_path / "a.fits").write_bytes(fits_bytes([("E", t), ("E", q), ("E", u)], nside, "RING", 1024, extra_hdu=True))
    for i, want in enumerate((t, q, u)):
        got, ns = ref_read(ref, tmp_path / "a.fits", i)
        assert ns == nside and np.array_equal(got, want)
    d = rng.standard_normal(npix)
    (tmp_path / "b.fits").write_bytes(fits_bytes([("E", t), ("D", d)], nside, "RING", 1))
    got, _ = ref_read(ref, tmp_path / "b.fits", 1)
    assert np.array_equal(got, d.astype(np.float32))
    (tmp_path / "c.fits").write_bytes(fits_bytes([("E", t)], nside, "NESTED", 1024))
    got, _ = ref_read(ref, tmp_path / "c.fits")
    ring = np.empty_like(t)
    ring[oracle.nest2ring(nside, np.arange(npix))] = t
    assert np.array_equal(got, ring)
