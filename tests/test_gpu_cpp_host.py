"""Compiles tests/cpp/test_host_mirror.cpp (g++) against librest_b200.so + the oracle and runs it on the GPU box:
the C++ host mirror (rest_tensors_b200/host/rest_tensors.hpp) over the C ABI, no Python in the call path."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
    from oracle.api import build as build_oracle
    build_oracle()
    exe = os.path.join(tmp, "test_host_mirror")
    lib_dir = os.path.join(ROOT, "rest_tensors_b200")
    orc_dir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", exe,
                           f"-L{lib_dir}", "-lrest_b200", f"-L{orc_dir}", "-lrest_oracle",
                           f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{orc_dir}"])
    return exe


def test_cpp_host_mirror_compiles_and_links(tmp_path):
    """CPU tier: the header compiles and links against the shared library (no GPU call is made)."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    assert os.path.exists(_build(str(tmp_path)))


@pytest.mark.gpu
def test_cpp_host_mirror_runs(tmp_path):
    exe = _build(str(tmp_path))
    from oracle.api import find_openblas
    env = dict(os.environ)
    blas = find_openblas()
    if blas:
        env["OPENBLAS_PATH"] = blas           # lets the C++ test compare lapack_dsyev with LAPACK's dsyev too
    out = subprocess.run([exe], capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0 and "CPP_HOST_MIRROR_OK" in out.stdout, out.stdout + out.stderr
