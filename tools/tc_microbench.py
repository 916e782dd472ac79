"""Micro-benchmark of the tensor-core GEMM kernel alone (conv1 / conv2+GN / plain shapes), used for bottleneck
experiments: TAG_TC_DEBUG=1 (no epilogue stores), 2 (no weight loads), 4 (no activation loads), TAG_TC_PAIR=0/1."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tag_b200 as tb
from tag_b200 import _lib

DEV = "cuda:0"
lib = _lib.load()
h = tb.scoring.util_handle(DEV)
s = torch.cuda.current_stream().cuda_stream


def bench(name, M, N, K, taps, dil, T, act, res, gn, out32=False, reps=20):
    g = torch.Generator(device=DEV).manual_seed(1)
    A = torch.randn(M, K, device=DEV, generator=g).half()
    W = (torch.randn(N, taps * K, device=DEV, generator=g) / math.sqrt(K * taps)).half()
    R16 = torch.randn(M, N, device=DEV, generator=g).half() if res == 16 else None
    R32 = torch.randn(M, N, device=DEV, generator=g) if res == 32 else None
    C16 = None if out32 else torch.empty(M, N, device=DEV, dtype=torch.float16)
    C32 = torch.empty(M, N, device=DEV) if out32 else None
    gam = torch.ones(N, device=DEV) if gn else None
    bet = torch.zeros(N, device=DEV) if gn else None
    def run():
        rc = lib.tag_debug_gemm_tc(h, A.data_ptr(), K, W.data_ptr(), M, N, K, taps, dil, T, None, _lib.ptr(R16), _lib.ptr(R32),
                                   _lib.ptr(C16), _lib.ptr(C32), act, _lib.ptr(gam), _lib.ptr(bet), None, None, s)
        _lib.check(h, rc, name)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * M * N * K * taps
    print(f"{name:28s} M={M} N={N} K={K} taps={taps}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)


W = 12500
print("env: PAIR=%s DEBUG=%s" % (os.environ.get("TAG_TC_PAIR", "1"), os.environ.get("TAG_TC_DEBUG", "0")))
bench("conv1 (gelu)", W * 32, 256, 256, 5, 2, 32, 1, 0, False)
bench("conv2 (res+gelu)", W * 32, 256, 256, 5, 2, 32, 1, 16, False)
bench("conv2+GN", W * 32, 256, 256, 5, 2, 32, 1, 16, True)
bench("conv1 no act", W * 32, 256, 256, 5, 2, 32, 0, 0, False)
bench("proj K=256", W * 32, 256, 256, 1, 1, 32, 0, 0, False)
bench("qkv N=768", W * 33, 768, 256, 1, 1, 1, 0, 0, False)
bench("ffn1 N=1024 relu", W * 33, 1024, 256, 1, 1, 1, 2, 0, False)
bench("ffn2 K=1024 res32->f32", W * 33, 256, 1024, 1, 1, 1, 0, 32, False, out32=True)
bench("outproj K=256 res32->f32", W * 33, 256, 256, 1, 1, 1, 0, 32, False, out32=True)
bench("big K=4096", W * 32, 256, 4096, 1, 1, 32, 0, 0, False)
