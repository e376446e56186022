#!/usr/bin/env python
"""Device-resident kernel bench (GPU box): tmm_device_gemm vs cuBLAS (torch.matmul) for double / complex<double>.
Development tool; cuBLAS is the comparator only."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import tiled_mm_b200 as tmm

def bench(dtype, tt, m, n, k, beta=0.0, reps=5):
    tdt = torch.float64 if dtype == np.float64 else torch.complex128
    ta, tb = tt
    ar, ac = (m, k) if ta == "N" else (k, m)
    br, bc = (k, n) if tb == "N" else (n, k)
    g = torch.Generator(device="cuda").manual_seed(1)
    def rnd(count):
        r = torch.rand(count, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
        return r if dtype == np.float64 else torch.complex(r, torch.rand(count, dtype=torch.float64, device="cuda", generator=g) * 2 - 1)
    A, B, C = rnd(ar * ac), rnd(br * bc), rnd(m * n)
    def opm(x, rows, cols, t):  # column-major rows x cols storage -> torch matrix op(x)
        mat = x.view(cols, rows).t()
        return mat if t == "N" else (mat.t() if t == "T" else mat.t().conj())
    Am, Bm = opm(A, ar, ac, ta), opm(B, br, bc, tb)
    ref = (Am @ Bm + beta * C.view(n, m).t())
    Cw = C.clone()
    st = torch.cuda.current_stream()
    tmm.device_gemm(dtype, ta, tb, m, n, k, 1.0, A.data_ptr(), ar, B.data_ptr(), br, beta, Cw.data_ptr(), m, stream=st.cuda_stream)
    err = float((Cw.view(n, m).t() - ref).abs().max()) / k
    def timeit(fn):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t_ours = timeit(lambda: tmm.device_gemm(dtype, ta, tb, m, n, k, 1.0, A.data_ptr(), ar, B.data_ptr(), br, beta, Cw.data_ptr(), m, stream=st.cuda_stream))
    t_cub = timeit(lambda: torch.matmul(Am, Bm))
    fl = (2.0 if dtype == np.float64 else 8.0) * m * n * k
    print(f"{np.dtype(dtype).name:10s} {tt} {m:6d} {n:6d} {k:6d} beta={beta}: ours {t_ours:8.3f} ms {fl/t_ours*1e-9:6.2f} TF | cuBLAS {t_cub:8.3f} ms {fl/t_cub*1e-9:6.2f} TF | ratio {t_cub/t_ours:.3f} | err/k {err:.2e}", flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "d"):
        for tt in ("NN", "TN", "NT", "TT"):
            bench(np.float64, tt, 10000, 10000, 10000)
        bench(np.float64, "NN", 10000, 4800, 512, beta=1.0)
        bench(np.float64, "NN", 10000, 4800, 2048, beta=1.0)
        bench(np.float64, "NN", 10000, 2048, 10000)
    if which == "z1":
        bench(np.complex128, "NN", 6000, 6000, 6000, reps=2)
    if which in ("all", "z"):
        for tt in ("NN", "CN", "NC", "TT"):
            bench(np.complex128, tt, 6000, 6000, 6000)
        bench(np.complex128, "NN", 6000, 6000, 512, beta=1.0)
        bench(np.complex128, "CN", 10000, 5000, 2048, beta=1.0)
