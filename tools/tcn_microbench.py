"""Fused TemporalConvBlock kernel against the two GEMM launches it replaces (conv1 + GELU, conv2 + residual + GELU + GroupNorm), alone,
on one encoder pass worth of rows (12,500 windows x 32 frames)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tag_b200 as tb
from tag_b200 import _lib

DEV = "cuda:0"
lib = _lib.load()
h = tb.scoring.util_handle(DEV)
W_, T, N, K, taps = 12500, 32, 256, 256, 5
M = W_ * T
g = torch.Generator(device=DEV).manual_seed(3)
x = torch.randn(M, K, device=DEV, generator=g).half()
W1 = (torch.randn(N, taps * K, device=DEV, generator=g) / math.sqrt(K * taps)).half()
W2 = (torch.randn(N, taps * K, device=DEV, generator=g) / math.sqrt(K * taps)).half()
gamma, beta = torch.ones(N, device=DEV), torch.zeros(N, device=DEV)
y1 = torch.empty_like(x)
s = torch.cuda.current_stream().cuda_stream


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for dil in (1, 2, 4, 8):
    buf = x.clone()

    def two():
        _lib.check(h, lib.tag_debug_gemm_tc(h, buf.data_ptr(), K, W1.data_ptr(), M, N, K, taps, dil, T, None, None, None, y1.data_ptr(), None, 1,
                                            None, None, None, None, s), "conv1")
        _lib.check(h, lib.tag_debug_gemm_tc(h, y1.data_ptr(), K, W2.data_ptr(), M, N, K, taps, dil, T, None, buf.data_ptr(), None, buf.data_ptr(),
                                            None, 1, gamma.data_ptr(), beta.data_ptr(), None, None, s), "conv2")

    def one():
        _lib.check(h, lib.tag_debug_tcn_block(h, buf.data_ptr(), M, T, dil, W1.data_ptr(), W2.data_ptr(), gamma.data_ptr(), beta.data_ptr(), s), "block")

    t2 = timed(two)
    fl = 2 * 2.0 * M * N * K * taps
    line = f"dil {dil}: two kernels {t2:7.1f} us ({fl / t2 / 1e6:6.1f} TFLOP/s)"
    if lib.tag_debug_tcn_block(h, buf.data_ptr(), M, T, dil, W1.data_ptr(), W2.data_ptr(), gamma.data_ptr(), beta.data_ptr(), s) == 0:
        t1 = timed(one)
        line += f"   fused {t1:7.1f} us ({fl / t1 / 1e6:6.1f} TFLOP/s)   {t2 / t1:.3f}x"
    else:
        line += "   fused: not supported (two-kernel path)"
    print(line)
