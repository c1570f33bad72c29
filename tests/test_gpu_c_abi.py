"""The drop-in boundary from a host that is neither Python nor PyTorch: tests/c/index_smoke.c is
compiled with gcc against include/ragarc_b200.h, linked with libragarc_b200.so, and run.
CPU part: the header is strict C99 and the program links and fails with RAGARC_ERR_CUDA (no
fallback).  GPU part: the program's own brute-force checks of add / search / remove pass."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "index_smoke.c")
LIBDIR = os.path.join(ROOT, "rag_arc_b200")


def _build(tmp_path):
    from rag_arc_b200 import _native  # noqa: F401 - makes sure the library has been built
    exe = str(tmp_path / "index_smoke")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    SRC, "-o", exe, "-L", LIBDIR, "-lragarc_b200", f"-Wl,-rpath,{LIBDIR}", "-lm"],
                   check=True, capture_output=True, text=True)
    return exe


def test_header_is_strict_c99_and_program_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 2, res.stderr
    assert "ragarc_index_create" in res.stderr and "failed (2)" in res.stderr     # RAGARC_ERR_CUDA


@pytest.mark.gpu
def test_plain_c_host_add_search_remove(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "index_smoke: ok" in res.stdout
