"""tcgen05 kind::tf32 GEMM probe: correctness of the four operand-order combinations against torch (and, if the
MN-major descriptor strides were guessed wrong, which (LBO, SBO) pair is right), then timing against the mma.sync TF32
kernel and the fp32 SIMT kernel on training-step shapes."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arreau_b200 import _lib  # noqa: E402

dev = torch.device("cuda")
lib = _lib.load()
lib.arreau_debug_set_tf32_gemm.argtypes = [C.c_int, C.c_int, C.c_int]
partial = torch.empty(8 << 20, device=dev)
s = torch.cuda.current_stream().cuda_stream


def run(ak, bk, A, B, Cm, M, N, K, tf32, alpha=1.0, bias=None, acc=0):
    _lib.call("arreau_sgemm", ak | (2 * tf32), bk, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), Cm.data_ptr(), N, M, N, K,
              C.c_float(alpha), None if bias is None else bias.data_ptr(), acc, partial.data_ptr(), partial.numel(), s)


def check(M, N, K, ak, bk):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((M, K) if ak else (K, M), generator=g).to(dev)
    B = torch.randn((N, K) if bk else (K, N), generator=g).to(dev)
    ref = (A.double() if ak else A.double().T) @ (B.double().T if bk else B.double())
    Cm = torch.zeros(M, N, device=dev)
    run(ak, bk, A, B, Cm, M, N, K, 1)
    torch.cuda.synchronize()
    return float((Cm.double() - ref).abs().max() / ref.abs().max())


ok = True
for lbo, sbo in ((512, 2048), (2048, 512), (512, 512), (128, 2048)):
    lib.arreau_debug_set_tf32_gemm(0, lbo, sbo)
    errs = {(ak, bk): check(256, 256, 512, ak, bk) for ak in (1, 0) for bk in (1, 0)}
    print(f"MN-major LBO={lbo} SBO={sbo}: rel err (ak,bk) -> {{{', '.join(f'{k}: {v:.2e}' for k, v in errs.items())}}}", flush=True)
    if all(v < 2e-3 for v in errs.values()):
        break
else:
    ok = False
    print("NO descriptor setting gave correct MN-major results")
if ok:
    for (M, N, K, ak, bk) in ((300, 128, 96, 1, 1), (128, 256, 20000, 0, 0), (1000, 512, 128, 1, 0), (128, 16, 256, 0, 0), (77, 128, 96, 1, 0),
                              (512, 128, 5000, 0, 0), (284928, 256, 128, 1, 1), (128, 256, 284928, 0, 0), (284928, 128, 256, 1, 0)):
        print(f"  M={M} N={N} K={K} ak={ak} bk={bk}: rel err {check(M, N, K, ak, bk):.2e}", flush=True)
    # timing on the training step's dominant shapes (C5: E*O = 285k rows)
    shapes = [("fwd/recompute z2 = a1 W2^T", 284928, 256, 128, 1, 1), ("input grad da1 = dz2 W2", 284928, 128, 256, 1, 0),
              ("weight grad dW2 = dz2^T a1", 256, 128, 284928, 0, 0), ("node MLP 35.6k x 512 x 128", 35616, 512, 128, 1, 1)]
    for name, M, N, K, ak, bk in shapes:
        A = torch.randn((M, K) if ak else (K, M), device=dev)
        B = torch.randn((N, K) if bk else (K, N), device=dev)
        Cm = torch.zeros(M, N, device=dev)
        res = []
        for label, tf32, legacy in (("fp32 SIMT", 0, 0), ("mma.sync TF32", 1, 1), ("tcgen05 TF32", 1, 0)):
            lib.arreau_debug_set_tf32_gemm(legacy, 0, 0)
            for _ in range(2):
                run(ak, bk, A, B, Cm, M, N, K, tf32)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run(ak, bk, A, B, Cm, M, N, K, tf32)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            res.append(f"{label} {ms * 1e3:.0f} us ({2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s)")
        print(f"{name}: " + ", ".join(res), flush=True)
lib.arreau_debug_set_tf32_gemm(0, 0, 0)
