"""The C-ABI library loads and exports every symbol include/gh_cuda.h declares (no compute, no GPU)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from crime_b200 import abi

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "gh_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gh_cuda_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(abi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.gh_cuda_version()


def test_params_struct_layout_matches_header_order():
    text = (ROOT / "include" / "gh_cuda.h").read_text()
    start = text.index("typedef struct gh_cuda_params {") + len("typedef struct gh_cuda_params {")
    body = text[start:text.index("} gh_cuda_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*(\[\d+\])?\s*$", part.strip())
            names.append(m.group(1))
    assert names == [n for n, _ in abi.GhCudaParams._fields_]


def test_create_fails_loudly_without_gpu_or_with_bad_params(tables_nu64):
    import torch
    from crime_b200.gethi import params_from_tables
    lib = abi.load_library()
    ctx = C.c_void_p()
    bad = abi.GhCudaParams()
    assert lib.gh_cuda_create(C.byref(bad), 0, 1, None, 0, C.byref(ctx)) != 0
    assert b"n_grid" in lib.gh_cuda_last_error()
    p = params_from_tables(tables_nu64, n_grid=47, n_side=8)  # odd: the half-complex layout needs an even grid
    assert lib.gh_cuda_create(C.byref(p), 0, 1, None, 0, C.byref(ctx)) != 0
    assert b"unsupported" in lib.gh_cuda_last_error()
    p = params_from_tables(tables_nu64, n_grid=48, n_side=8)  # general-length FFT: one rank only
    assert lib.gh_cuda_create(C.byref(p), 0, 2, b"\0" * 128, 0, C.byref(ctx)) != 0
    assert b"unsupported on 2 ranks" in lib.gh_cuda_last_error()
    p = params_from_tables(tables_nu64, n_grid=64, n_side=8)
    assert lib.gh_cuda_create(C.byref(p), 0, 3, None, 0, C.byref(ctx)) != 0  # 3 ranks: not a power of two
    if not torch.cuda.is_available():
        # the product must not silently fall back to a CPU path
        assert lib.gh_cuda_create(C.byref(p), 0, 1, None, 0, C.byref(ctx)) != 0
        assert b"no CPU fallback" in lib.gh_cuda_last_error()
        from crime_b200 import GetHI, GetHIError
        with pytest.raises(GetHIError):
            GetHI(p)


def test_missing_library_raises(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        abi.load_library(tmp_path / "libgh_cuda.so")


def test_slab_arithmetic():
    from crime_b200 import slab
    assert slab.slab_bounds(512, 1, 0) == (512, 0)
    assert slab.slab_bounds(2048, 8, 3) == (256, 768)
    assert slab.shell_bounds(150, 8, 7) == (17, 133, 152)
    assert slab.shell_bounds(150, 8, 0) == (19, 0, 152)
    assert slab.shell_bounds(64, 1, 0) == (64, 0, 64)
    assert sum(slab.shell_bounds(150, 8, r)[0] for r in range(8)) == 150
    with pytest.raises(ValueError):
        slab.slab_bounds(512, 3, 0)
    # every (z_local, ky, kx) lands in a distinct slot of the receive buffer
    n, p = 8, 2
    seen = {slab.received_index(n, p, z, ky, kx) for z in range(n // p) for ky in range(n) for kx in range(n // 2 + 1)}
    assert len(seen) == (n // p) * n * (n // 2 + 1) and max(seen) == len(seen) - 1
