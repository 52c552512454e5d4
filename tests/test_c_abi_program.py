"""The C ABI from plain C (no ctypes, no Python objects across the boundary): tests/c_abi/test_c_abi.c includes
include/nagp.h, links libnagp.so and runs one nagp_forecast_with_nowcasts call, checked against closed forms.
CPU: the program compiles and links against the built library (every symbol it uses resolves). GPU: it runs."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "test_c_abi.c")
LIBDIR = os.path.join(ROOT, "nowcastautogp_b200")


def _build(tmp_path):
    from nowcastautogp_b200 import build
    build.build()
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    exe = os.path.join(str(tmp_path), "test_c_abi")
    cmd = [cc, "-O1", "-std=c11", "-D_GNU_SOURCE", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-lnagp", f"-Wl,-rpath,{LIBDIR}", "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_program_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_c_program_runs(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "c-abi ok" in res.stdout
