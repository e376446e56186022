"""the CMake package under the reference's names.  Configure + build + install the library with CMake, then a
downstream project does find_package(Tiled-MM) / Tiled-MM::Tiled-MM exactly as it would against the reference."""
import os
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _run(cmd, **kw):
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, **kw)
    assert r.returncode == 0, (" ".join(map(str, cmd)), r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout


@pytest.mark.skipif(shutil.which("cmake") is None or shutil.which("nvcc") is None, reason="needs cmake and nvcc")
def test_find_package_tiled_mm(tmp_path):
    build, prefix, cbuild = tmp_path / "build", tmp_path / "prefix", tmp_path / "consumer"
    gen = ["-G", "Ninja"] if shutil.which("ninja") else []
    _run(["cmake", "-S", str(ROOT), "-B", str(build), *gen, f"-DCMAKE_INSTALL_PREFIX={prefix}"])
    _run(["cmake", "--build", str(build), "-j8"])
    _run(["cmake", "--install", str(build)])
    # install layout of the reference (CMakeLists.txt:44-63): lib/, lib/cmake/Tiled-MM/, include/Tiled-MM/*.hpp
    libdir = next(p for p in (prefix / "lib", prefix / "lib64") if p.exists())
    assert (libdir / "cmake" / "Tiled-MM" / "Tiled-MMConfig.cmake").exists() and (libdir / "cmake" / "Tiled-MM" / "Tiled-MMTargets.cmake").exists()
    for h in ("tiled_mm.hpp", "mm_handle.hpp", "util.hpp", "device_vector.hpp", "device_buffer.hpp", "gpu_blas_api.hpp", "gpu_runtime_api.hpp", "gpu_blas_handle.hpp"):
        assert (prefix / "include" / "Tiled-MM" / h).exists(), h
    # the same mangled drop-in symbols and C ABI as the Makefile build
    syms = _run(["nm", "-D", "--defined-only", str(libdir / "libTiled-MM.so")])
    assert "_ZN3gpu4gemmIdEEvRNS_9mm_handleIT_EEcciiiS2_PS2_iS5_iS2_S5_ibb" in syms and " tmm_gemm" in syms
    assert (build / "test-multiply").exists() and (build / "multiply").exists()
    tests = _run(["ctest", "-N"], cwd=str(build))
    for name in ("square-small", "square-large", "non-square-small", "non-square-large"):  # reference tests/CMakeLists.txt:12-15
        assert name in tests
    # downstream project
    _run(["cmake", "-S", str(ROOT / "tests" / "cmake_consumer"), "-B", str(cbuild), *gen, f"-DCMAKE_PREFIX_PATH={prefix}"])
    _run(["cmake", "--build", str(cbuild)])
    assert "libTiled-MM.so" in _run(["ldd", str(cbuild / "consumer")])
    assert subprocess.run([str(cbuild / "consumer")]).returncode == 0


def test_rocm_backend_is_refused(tmp_path):
    if shutil.which("cmake") is None:
        pytest.skip("needs cmake")
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(["cmake", "-S", str(ROOT), "-B", str(tmp_path / "b"), "-DTILEDMM_GPU_BACKEND=ROCM"], capture_output=True, text=True, env=env)
    assert r.returncode != 0 and "must be CUDA" in r.stderr
