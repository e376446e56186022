"""Generate golden input/output vectors by running the REFERENCE ITSELF.

  python tests/golden/make_golden.py            # here: unmodified reference scheduler over the CUDA/cuBLAS emulation
  python tests/golden/make_golden.py --cublas   # on the GPU box: unmodified reference + real cuBLAS (gpurun), writes
                                                # into gpurun_out/golden/, copied into tests/golden/ afterwards

Inputs come from the reference test's own generator (mt19937(42), ints 0..9; tests/test-multiply.cpp:58-66) when
`seed42` is set - then only the outputs are stored - or are stored explicitly (small random cases).
"""
import argparse
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import _util  # noqa: E402


def run(ref, oracle, name, dtype, tt, m, n, k, tiles, alpha, beta, pad=(0, 0, 0), seed42=True, exact=True, outdir=HERE):
    ta, tb = tt
    ar, ac = _util.stored_shape(ta, m, k)
    br, bc = _util.stored_shape(tb, k, n)
    lda, ldb, ldc = ar + pad[0], br + pad[1], m + pad[2]
    if seed42:
        a, b, c = oracle.fixture_abc(dtype, lda * ac, ldb * bc, ldc * n)
    else:
        rng = np.random.default_rng(2024)
        a, b, c = (_util.random_matrix(rng, dtype, x) for x in (lda * ac, ldb * bc, ldc * n))
    ctx = ref.context(dtype, 2, *tiles)
    out = c.copy()
    ctx.gemm(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, out, ldc, pin=False, copy_c_back=True)
    ctx.close()
    payload = dict(trans=np.array(tt), mnk=np.array([m, n, k]), lds=np.array([lda, ldb, ldc]), tiles=np.array(tiles),
                   alpha=np.array(alpha, dtype=dtype), beta=np.array(beta, dtype=dtype), c_out=out, seed42=np.array(seed42), exact=np.array(exact))
    if not seed42:
        payload.update(a=a, b=b, c_in=c)
    np.savez_compressed(outdir / f"{name}.npz", **payload)
    print("wrote", name, out.dtype, out.shape)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cublas", action="store_true")
    args = ap.parse_args()
    oracle = _util.Oracle()
    ref = _util.Reference(cpu=not args.cublas)
    if not args.cublas:
        run(ref, oracle, "ci_50x200x21_tile4", np.float64, "NN", 50, 200, 21, (4, 4, 4), 1.0, 0.0)       # ci/daint-alps.yml:51
        run(ref, oracle, "ci_5x2x2_tile4", np.float64, "NN", 5, 2, 2, (4, 4, 4), 1.0, 0.0)               # ci/daint-alps.yml:60
        run(ref, oracle, "c1_200cubed_beta1", np.float64, "NN", 200, 200, 200, (5000, 5000, 5000), 1.0, 1.0)   # config C1 shape class
        run(ref, oracle, "tn_ld_z", np.complex128, "CN", 33, 41, 48, (16, 16, 12), 1 - 1j, 0.5 + 0j, pad=(3, 1, 2))
        run(ref, oracle, "nt_ld_s", np.float32, "NT", 64, 31, 40, (20, 16, 10), 2.0, -1.0, pad=(1, 0, 5))
    else:
        out = HERE.parent.parent / "gpurun_out" / "golden"
        out.mkdir(parents=True, exist_ok=True)
        # the reference's registered ctest case and BASELINE config C1, through real cuBLAS
        run(ref, oracle, "c1_1000cubed_beta1_cublas", np.float64, "NN", 1000, 1000, 1000, (5000, 5000, 5000), 1.0, 1.0, outdir=out)
        run(ref, oracle, "ci_50x200x21_tile4_cublas", np.float64, "NN", 50, 200, 21, (4, 4, 4), 1.0, 0.0, outdir=out)
        run(ref, oracle, "rand_tn_d_cublas", np.float64, "TN", 257, 131, 300, (100, 64, 100), 1.5, -0.5, pad=(1, 2, 3), seed42=False, exact=False, outdir=out)
        run(ref, oracle, "rand_cn_z_cublas", np.complex128, "CN", 129, 65, 160, (64, 64, 80), 1 - 2j, 0.5 + 1j, pad=(0, 2, 1), seed42=False, exact=False, outdir=out)
        run(ref, oracle, "rand_nt_s_cublas", np.float32, "NT", 200, 100, 150, (64, 64, 50), 1.0, 1.0, seed42=False, exact=False, outdir=out)
        run(ref, oracle, "rand_nc_c_cublas", np.complex64, "NC", 100, 120, 90, (64, 64, 30), 1 + 1j, 0j, seed42=False, exact=False, outdir=out)


if __name__ == "__main__":
    main()
