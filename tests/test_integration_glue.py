"""INTEGRATION.md shows the C file a maintainer of the reference would add.  Where the reference tree is present
(the build container), that code block is compiled against the reference's own unmodified headers (plus the
stand-in GSL/FFTW/chealpix headers the oracle build uses) and include/gh_cuda.h, in both frequency-table
personalities: the binding in the document is real code, not a sketch."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/src")


@pytest.mark.skipif(not (REF / "common_gh.h").exists(), reason="reference tree not present on this box")
@pytest.mark.parametrize("defs", [["-D_IRREGULAR_NUTABLE"], []])
def test_glue_in_integration_md_compiles_against_the_reference_headers(tmp_path, defs):
    text = (ROOT / "INTEGRATION.md").read_text()
    blocks = re.findall(r"```c\n(.*?)```", text, flags=re.S)
    glue = [b for b in blocks if "gh_cuda_glue.c" in b]
    assert len(glue) == 1
    # oracle/gh_cuda_glue.c (what oracle/Makefile links into _ref/GetHI_gpu) is the document's text
    assert (ROOT / "oracle" / "gh_cuda_glue.c").read_text() == glue[0]
    src = tmp_path / "gh_cuda_glue.c"
    src.write_text(glue[0])
    cmd = ["gcc", "-std=gnu99", "-c", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
           "-D_LONGIDS", "-D_DEBUG", "-D_HAVE_OMP", "-D_SPREC", *defs, f"-I{REF}", f"-I{ROOT / 'oracle' / 'shim'}",
           f"-I{ROOT / 'include'}", str(src), "-o", str(tmp_path / "glue.o")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    syms = subprocess.run(["nm", str(tmp_path / "glue.o")], capture_output=True, text=True).stdout
    for f in ("init_fftw", "create_d_and_vr_fields", "get_HI", "mk_T_maps", "end_fftw"):
        assert re.search(rf" T {f}\b", syms), f"{f} not defined by the glue"
    for f in ("gh_cuda_create", "gh_cuda_create_d_and_vr_fields", "gh_cuda_get_HI", "gh_cuda_mk_T_maps", "gh_cuda_destroy"):
        assert re.search(rf" U {f}\b", syms), f"{f} not referenced by the glue"


@pytest.mark.skipif(not (REF / "main_gh.c").exists(), reason="reference tree not present on this box")
def test_reference_driver_links_against_libgh_cuda_and_fails_loudly_without_a_gpu(tmp_path):
    """oracle/_ref/GetHI_gpu = the reference's own main_gh.c / io_gh.c / cosmo.c ... + the glue + libgh_cuda.so.
    Without a GPU it must get as far as init_fftw (the reference's parameter reader and cosmology run) and then
    stop through the reference's own report_error with the library's message."""
    import sys
    import torch
    sys.path.insert(0, str(ROOT))
    from oracle.binding import write_nutable, write_param_file
    exe = ROOT / "oracle" / "_ref" / "GetHI_gpu"
    if not exe.exists():
        pytest.skip("oracle/_ref/GetHI_gpu not built")
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run itself is covered by the gpu-marked test")
    write_nutable(tmp_path / "nu.txt", 8)
    write_param_file(tmp_path / "p.ini", n_grid=32, n_side=16, nutable=tmp_path / "nu.txt",
                     pk_file=ROOT / "data" / "Pk_synth.dat", prefix=tmp_path / "out")
    r = subprocess.run([str(exe), str(tmp_path / "p.ini")], capture_output=True, text=True, timeout=120, cwd=tmp_path)
    out = r.stdout + r.stderr
    assert r.returncode != 0
    assert "Reading P_k from file" in out                      # the reference's own reader and cosmology ran
    assert "Fatal" in out and "no CPU fallback" in out         # ... and its own report_error carried our message
    assert not list(tmp_path.glob("out_*.fits"))
