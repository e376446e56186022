"""CPU suite: the C-ABI library loads without a GPU, exports every symbol include/tiled_mm_b200.h declares, and
refuses to compute without a device (no CPU fallback)."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(tmm):
    lib = tmm.load_library()
    declared = tmm.declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tiled_mm_b200.h but not exported"
    # and nothing is declared twice with a different spelling in the python mirror
    for name in ("tmm_context_create", "tmm_gemm", "tmm_context_device_c", "tmm_malloc_pinned", "tmm_device_gemm"):
        assert name in declared


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C with no torch / C++ types."""
    src = tmp_path / "t.c"
    src.write_text('#include "tiled_mm_b200.h"\nint main(void){ tmm_call_stats s; (void)s; return TMM_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_cpp_dropin_symbols_exported(tmm):
    """Same mangled names as the reference library's explicit instantiations (tiled_mm.cpp:626-668, mm_handle.cpp:167-170)."""
    out = subprocess.run(["nm", "-D", "--defined-only", str(tmm.LIB_PATH)], check=True, capture_output=True, text=True).stdout
    for sym in [
        "_ZN3gpu4gemmIdEEvRNS_9mm_handleIT_EEcciiiS2_PS2_iS5_iS2_S5_ibb",
        "_ZN3gpu4gemmIfEEvRNS_9mm_handleIT_EEcciiiS2_PS2_iS5_iS2_S5_ibb",
        "_ZN3gpu4gemmISt7complexIdEEEvRNS_9mm_handleIT_EEcciiiS4_PS4_iS7_iS4_S7_ibb",
        "_ZN3gpu4gemmISt7complexIfEEEvRNS_9mm_handleIT_EEcciiiS4_PS4_iS7_iS4_S7_ibb",
        "_ZN3gpu9mm_handleIdEC1Eiiii",
        "_ZN3gpu9mm_handleIdE24get_full_device_buffer_cEv",
        "_ZN3gpu9mm_handleIdE18optimal_tile_sizesEiii",
        "_ZN3gpu18get_blas_operationEc",
        "_ZN3gpu9mm_handleIdE19get_device_buffer_aEv",
        "_ZN3gpu9mm_handleIfE19get_device_buffer_cEv",
        "_ZN3gpu9mm_handleISt7complexIdEE21set_streams_and_tilesEiiii",
        "_ZN3gpu9mm_handleIdE14set_full_sizesEiii",
    ]:
        assert sym in out, sym


def test_cpp_dropin_headers_compile(tmp_path):
    """A reference-style caller (tests/test-multiply.cpp shape) compiles against include/Tiled-MM unchanged."""
    src = tmp_path / "caller.cpp"
    src.write_text(
        "#include <Tiled-MM/tiled_mm.hpp>\n#include <Tiled-MM/device_vector.hpp>\n#include <Tiled-MM/util.hpp>\n#include <Tiled-MM/gpu_blas_handle.hpp>\n"
        "#include <Tiled-MM/gpu_blas_api.hpp>\n#include <Tiled-MM/gpu_runtime_api.hpp>\n#include <Tiled-MM/device_buffer.hpp>\n#include <Tiled-MM/gpu_context.hpp>\n#include <Tiled-MM/tiled_matrix.hpp>\n#include <Tiled-MM/tile_coord.hpp>\n"
        "int run(int m, int n, int k) {\n"
        "  auto a = gpu::malloc_pinned<double>(size_t(m) * k, 1); auto b = gpu::malloc_pinned<double>(size_t(k) * n, 1);\n"
        "  auto c = gpu::malloc_pinned<double>(size_t(m) * n, 0);\n"
        "  auto ctx = gpu::make_context<double>(2, 5000, 5000, 5000);\n"
        "  gpu::gemm(*ctx, 'N', 'N', m, n, k, 1.0, a, m, b, k, 0.0, c, m, false, true);\n"
        "  gpu::gemm(*ctx, 'N', 'N', m, n, k, 1.0, a, m, b, k, 0.0, c, m, false, false);\n"
        "  gpu::copy_to_host(ctx->get_full_device_buffer_c().data(), c, size_t(m) * n);\n"
        "  auto z = gpu::make_context<std::complex<double>>();\n"
        "  // the rest of the handle surface (mm_handle.hpp:22-42): slabs, full C sizing, streams\n"
        "  ctx->set_num_streams(3); ctx->set_tile_sizes(100, 200, 300); ctx->set_tile_sizes(64); ctx->set_streams_and_tiles(2, 10, 20, 30);\n"
        "  ctx->set_full_sizes(m, n, k);\n"
        "  gpu::device_buffer<double>& ab = ctx->get_device_buffer_a(); gpu::tile_dim td = ab.get_tile_sizes();\n"
        "  double* slab1 = ctx->get_device_buffer_c().stream_buffer(1); double* base = ctx->get_device_buffer_b().data(); (void)slab1; (void)base;\n"
        "  gpu::gpu_context& gc = ctx->get_gpu_context(); cudaStream_t s0 = gc.get_stream(0); cudaStream_t rs = gc.get_result_stream().stream(); gpu::device_stream& ds = gc.get_device_stream(1); (void)s0; (void)rs; (void)ds;\n"
        "  gpu::blas_api::HandleType bh; gpu::blas_api::create(&bh); gpu::blas_api::set_stream(bh, s0); gpu::blas_api::destroy(bh);\n"
        "  gpu::tiled_matrix<double> tmx(a, m, k, m, gpu::tile_dim(7, 5)); gpu::tile_coord tc(1, 2);\n"
        "  int tq = tmx.num_tiles_row() + tmx.num_tiles_col() + tmx.tile_dimensions(tc).rows() + tmx.tile_offset(tc) + (tmx.tile_data(tc) != nullptr);\n"
        "  int t0 = tq + td.rows() + td.cols() + td.size() + std::get<2>(ctx->get_max_tile_sizes()) + gc.get_num_streams();\n"
        "  return (int)ctx->get_num_streams() + (int)std::get<0>(ctx->optimal_tile_sizes(m, n, k)) + (gpu::get_blas_operation('T') == gpu::blas_api::operation::Transpose) + t0;\n"
        "}\n")
    subprocess.run(["g++", "-std=c++14", "-Wall", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include", "-c", str(src), "-o", str(tmp_path / "c.o")],
                   check=True)


def test_tiling_view_header_matches_the_reference_rules(tmp_path, oracle):
    """include/Tiled-MM/tiled_matrix.hpp (kept for callers that include it): clamp, ceiling tile count, remainder tile, offsets - the
    rules of reference tiled_matrix.cpp:8-23,62-80, here with 64-bit offsets; tile counts cross-checked against the oracle's restatement."""
    src = tmp_path / "tm.cpp"
    src.write_text(
        "#include <Tiled-MM/tiled_matrix.hpp>\n#include <cstdio>\n#include <cstdlib>\n"
        "int main(int argc, char** argv) {\n"
        "  double buf[1]; int rows = atoi(argv[1]), cols = atoi(argv[2]), tr = atoi(argv[3]), tc = atoi(argv[4]);\n"
        "  gpu::tiled_matrix<double> t(buf, rows, cols, rows + 3, gpu::tile_dim(tr, tc));\n"
        "  gpu::tile_coord last(t.num_tiles_row() - 1, t.num_tiles_col() - 1);\n"
        "  printf(\"%d %d %d %d %zu\\n\", t.num_tiles_row(), t.num_tiles_col(), t.tile_dimensions(last).rows(), t.tile_dimensions(last).cols(), t.tile_offset64(last));\n"
        "}\n")
    exe = tmp_path / "tm"
    subprocess.run(["g++", "-std=c++14", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    for rows, cols, tr, tc in [(12345, 23456, 4115, 2932), (5, 2, 4, 4), (50, 200, 4, 4), (1000, 1000, 5000, 5000), (60000, 60000, 5000, 5000)]:
        out = subprocess.run([str(exe), str(rows), str(cols), str(tr), str(tc)], check=True, capture_output=True, text=True).stdout.split()
        ntr, ntc, lr, lc, off = (int(x) for x in out)
        ctr, ctc = min(tr, rows), min(tc, cols)
        assert ntr == oracle.lib.oracle_num_tiles(rows, ctr) and ntc == oracle.lib.oracle_num_tiles(cols, ctc)
        assert lr == rows - ctr * (ntr - 1) and lc == cols - ctc * (ntc - 1)
        assert off == (ntc - 1) * ctc * (rows + 3) + (ntr - 1) * ctr   # beyond 2^31 for the last case


def test_no_cpu_fallback_without_gpu(tmm):
    if tmm.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tmm.make_context(np.float64)
    with pytest.raises(ValueError):
        tmm.make_context(np.int32)


def test_optimal_tile_size_matches_oracle(tmm, oracle):
    for max_tile in (4, 7, 100, 5000):
        for dim in list(range(1, 60)) + [999, 1000, 1234, 4567, 1357, 5000, 5001, 10000, 12345, 23456, 67891]:
            assert tmm.optimal_tile_size(dim, max_tile) == oracle.lib.oracle_optimal_tile_size(dim, max_tile), (dim, max_tile)


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle (or any CPU GEMM)."""
    for f in list((ROOT / "tiled-mm_b200").rglob("*.py")) + list((ROOT / "tiled-mm_b200" / "csrc").glob("*.*")) + list((ROOT / "include").rglob("*.h*")):
        text = f.read_text(errors="ignore")
        assert not re.search(r"liboracle|oracle_gemm|oracle/|import _util", text), f
