"""tcgen05.mma kind::tf32 issue-rate probe (one CTA): time per MMA for N in {128, 256}."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from tensoflow_b200 import _lib
from tensoflow_b200._lib import check, ptr, stream_ptr
dev = torch.device('cuda:0')
lib = _lib.load()
for N, K in ((128, 32), (256, 32), (256, 56)):
    A = torch.randn(128, K, device=dev); B = torch.randn(N, K, device=dev); D = torch.empty(128, N, device=dev)
    for passes in (1, 3):
        res = {}
        for rep in (200, 2200):
            check(lib.tf_tc_probe(ptr(A), ptr(B), N, K, passes, rep, ptr(D), stream_ptr()), "probe")
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            check(lib.tf_tc_probe(ptr(A), ptr(B), N, K, passes, rep, ptr(D), stream_ptr()), "probe")
            b.record(); torch.cuda.synchronize()
            res[rep] = a.elapsed_time(b)
        n_mma = (2200 - 200) * (K // 8) * passes
        us = (res[2200] - res[200]) * 1e3 / n_mma
        print(f"N={N} K={K} passes={passes}: {us * 1e3:.1f} ns per MMA = {us * 1.965e3:.0f} cycles @1.965 GHz; "
              f"{128 * N * 8 * 2 / (us * 1e-6) / 1e12:.2f} TFLOP/s per SM -> x148 = {128 * N * 8 * 2 / (us * 1e-6) / 1e12 * 148:.0f} TFLOP/s")
