"""The C ABI exercised from plain C (tests/c_abi/test_capi.c): compiled with `cc` against include/odinn_b200.h and linked with
libodinn_b200.so -- no Python binding in between.  CPU: it compiles and links (every symbol it uses resolves).  GPU: it runs."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "test_capi.c")
LIBDIR = os.path.join(ROOT, "odinn.jl_b200", "lib")


def _build(tmp_path):
    import __graft_entry__ as ge

    ge.build_cuda()
    cc = shutil.which("cc") or shutil.which("gcc")
    exe = str(tmp_path / "test_capi")
    subprocess.run([cc, "-std=c11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-L", LIBDIR,
                    "-lodinn_b200", "-lm", "-Wl,-rpath," + LIBDIR, "-o", exe], check=True, capture_output=True, text=True)
    return exe


def test_c_program_compiles_and_links_against_the_abi(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_c_program_runs_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "test_capi: ok" in r.stdout
