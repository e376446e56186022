"""Command-line apps: this repository's multiply / test-multiply keep the reference miniapps' flags and report
format, and the reference's OWN app sources (examples/multiply.cpp, tests/test-multiply.cpp) compile unchanged against
include/Tiled-MM.  CPU: parsing, usage, loud failure without a GPU.  GPU: the reference's registered ctest cases."""
import os
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "bin"
REF = Path("/root/reference")


def _env():
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    return env


@pytest.fixture(scope="module")
def apps():
    subprocess.run(["make", "-C", str(ROOT / "tiled-mm_b200" / "csrc"), "-j4", "all"], check=True, capture_output=True, env=_env())
    r = subprocess.run(["make", "-C", str(ROOT / "apps"), "-j4", "all"], capture_output=True, text=True, env=_env())
    assert r.returncode == 0, r.stderr[-3000:]
    return BIN


def _run(path, *args, timeout=900):
    return subprocess.run([str(path), *map(str, args)], capture_output=True, text=True, timeout=timeout)


def test_usage_lists_the_reference_flags(apps):
    for app, extra in (("multiply", ["n_rep"]), ("test-multiply", [])):
        r = _run(apps / app, "--help")
        assert r.returncode == 0
        for flag in ["m_dim", "n_dim", "k_dim", "tile_m", "tile_n", "tile_k", "n_streams", "ld_a", "ld_b", "ld_c", "transpose", "alpha", "beta"] + extra:
            assert f"--{flag}" in r.stdout, (app, flag)


def test_bad_transpose_is_reported_like_the_reference(apps):
    # examples/multiply.cpp:83-91: "[ERROR]: --transpose option ..." on stdout, exit code 0
    r = _run(apps / "multiply", "-t", "NX")
    assert r.returncode == 0 and "[ERROR]: --transpose option" in r.stdout
    r = _run(apps / "test-multiply", "--bogus", "1")
    assert r.returncode == 2 and "unknown option" in r.stderr


@pytest.mark.skipif(not (REF / "tests" / "test-multiply.cpp").exists(), reason="reference sources not on this box")
def test_reference_apps_compile_unchanged(apps):
    """The reference's own miniapp and test app build against include/Tiled-MM + libtiledmm_b200.so without edits."""
    for name in ("ref-multiply", "ref-test-multiply"):
        assert (apps / name).exists()
        out = subprocess.run(["ldd", str(apps / name)], capture_output=True, text=True).stdout
        assert "libtiledmm_b200.so" in out and "cublas" not in out.lower()


def test_apps_fail_loudly_without_a_gpu(apps, tmm):
    if tmm.device_count() > 0:
        pytest.skip("a GPU is present")
    r = _run(apps / "test-multiply", "-m", 8, "-n", 8, "-k", 8)
    assert r.returncode != 0 and "GPU ERROR" in r.stderr


def test_cxxopts_shim_parses_like_the_reference_expects(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(
        '#include <cxxopts.hpp>\n#include <iostream>\nint main(int argc, char** argv) {\n'
        '  cxxopts::Options o("t", "d");\n'
        '  o.add_options()("m,m_dim", "rows", cxxopts::value<int>()->default_value("1000"))("tile_m", "tile", cxxopts::value<int>()->default_value("5000"))\n'
        '    ("t,transpose", "tr", cxxopts::value<std::string>()->default_value("NN"))("alpha", "a", cxxopts::value<double>()->default_value("1.0"));\n'
        '  auto r = o.parse(argc, argv);\n'
        '  if (r.count("help")) { std::cout << o.help(); return 0; }\n'
        '  std::cout << r["m_dim"].as<int>() << " " << r["tile_m"].as<int>() << " " << r["transpose"].as<std::string>() << " " << r["alpha"].as<double>() << "\\n";\n'
        '}\n')
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", "-I", str(ROOT / "tools" / "compat"), str(src), "-o", str(exe)], check=True, env=_env())
    assert _run(exe).stdout.split() == ["1000", "5000", "NN", "1"]
    assert _run(exe, "-m", "12", "--tile_m=7", "-t", "tn", "--alpha", "2.5").stdout.split() == ["12", "7", "tn", "2.5"]
    assert "--m_dim" in _run(exe, "--help").stdout


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def gpu_apps(gpu_tmm):
    """The binaries normally travel with the repository (built by __graft_entry__.build()); if they did not, build them on the box
    (same image: g++ is there; the reference's own sources are not, so bin/ref-* may legitimately be absent)."""
    if not (BIN / "test-multiply").exists() or not (BIN / "multiply").exists():
        r = subprocess.run(["make", "-C", str(ROOT / "apps"), "-j4", "all"], capture_output=True, text=True, env=_env())
        assert r.returncode == 0, r.stderr[-3000:]
    return BIN


CTEST_CASES = [  # reference tests/CMakeLists.txt:12-15 (the 10000^3 and 12345x23456x67891 cases are covered by test_gemm_gpu / bench)
    ("-m", 1000, "-n", 1000, "-k", 1000),
    ("-m", 1234, "-n", 4567, "-k", 1357),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CTEST_CASES)
def test_own_test_multiply_ctest_cases(gpu_apps, case):
    r = _run(BIN / "test-multiply", *case)
    assert r.returncode == 0 and "The result is CORRECT" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.gpu
@pytest.mark.parametrize("args", [
    ("--type", "z", "-t", "CT", "-m", 301, "-n", 403, "-k", 209, "--alpha", 2, "--beta", 1, "--ld_a", 211, "--ld_b", 405, "--ld_c", 307),
    ("--type", "s", "-t", "TN", "-m", 513, "-n", 300, "-k", 777, "--beta", 1, "--tile_m", 128, "--tile_n", 64, "--tile_k", 100),
    ("--type", "c", "-t", "NC", "-m", 129, "-n", 65, "-k", 300, "--beta", -1),
    ("--type", "d", "-t", "tt", "-m", 5, "-n", 2, "-k", 2, "--tile_m", 4, "--tile_n", 4, "--tile_k", 4),
])
def test_own_test_multiply_types_and_transposes(gpu_apps, args):
    r = _run(BIN / "test-multiply", *args)
    assert r.returncode == 0 and "The result is CORRECT" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CTEST_CASES)
def test_reference_test_app_passes_on_this_library(gpu_apps, case):
    """The reference's tests/test-multiply.cpp, compiled unchanged, run against this library (binary built where the reference
    sources exist; it travels with the repository).  NOT a parity test: the app's own oracle, compute_reference, calls blas_api::dgemm,
    which here is this library's device GEMM - so this checks that the drop-in headers and the scheduler agree with the one-shot device
    kernel on the reference's ctest cases; parity against cuBLAS / the oracle is what tests/test_gemm_gpu.py establishes."""
    exe = BIN / "ref-test-multiply"
    if not exe.exists():
        pytest.skip("bin/ref-test-multiply was not built (reference sources absent at build time)")
    r = _run(exe, *case)
    assert r.returncode == 0 and "The result is CORRECT" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.gpu
def test_multiply_report_format(gpu_apps):
    for exe in (BIN / "multiply", BIN / "ref-multiply"):
        if not exe.exists():
            continue
        r = _run(exe, "-m", 2000, "-n", 1500, "-k", 1000, "-r", 2, "-t", "NT")
        assert r.returncode == 0, r.stderr[-2000:]
        assert len(re.findall(r"-> Avg Time \[ms\] = ", r.stdout)) == 2 and len(re.findall(r"-> Throughput \[Gflops\] = ", r.stdout)) == 2
        assert " A = (2000, 1000)" in r.stdout and " B = (1500, 1000)" in r.stdout and " trans_b = T" in r.stdout
