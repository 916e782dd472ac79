"""Conv-layer micro-benchmark of the tensor-core GEMM (all four dilations, conv1 and conv2+GN forms).
TAG_TC_HALO=0 five shifted loads per chunk, 1 (default) halo tiles when the tap shift is swizzle-atom aligned."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tag_b200 as tb
from tag_b200 import _lib

DEV = "cuda:0"
lib = _lib.load()
h = tb.scoring.util_handle(DEV)
s = torch.cuda.current_stream().cuda_stream
W_, T, N, K = 12500, 32, 256, 256
M = W_ * T
g = torch.Generator(device=DEV).manual_seed(1)
A = torch.randn(M, K, device=DEV, generator=g).half()
Wt = (torch.randn(N, 5 * K, device=DEV, generator=g) / math.sqrt(5 * K)).half()
R16 = torch.randn(M, N, device=DEV, generator=g).half()
C16 = torch.empty(M, N, device=DEV, dtype=torch.float16)
gam, bet = torch.ones(N, device=DEV), torch.zeros(N, device=DEV)
print("env: HALO=%s ASTAGES=%s PAIR=%s DEBUG=%s" % (os.environ.get("TAG_TC_HALO", "0"), os.environ.get("TAG_TC_HALO_ASTAGES", "3"), os.environ.get("TAG_TC_PAIR", "1"), os.environ.get("TAG_TC_DEBUG", "0")))
if os.environ.get("TAG_CUBLAS", "0") != "0":
    # same-shape library comparator (cuBLAS through torch.matmul, fp16 operands, fp32 accumulate): the conv as an im2col GEMM
    # (K = 1280; the im2col matrix is NOT built by our kernel and is not charged here either) and the K = 4096 plain GEMM
    for Kc in (1280, 4096):
        Ac = torch.randn(M, Kc, device=DEV, generator=g).half()
        Bc = torch.randn(Kc, N, device=DEV, generator=g).half()
        for _ in range(3):
            torch.matmul(Ac, Bc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            torch.matmul(Ac, Bc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"cuBLAS fp16 [{M} x {Kc}] @ [{Kc} x {N}]: {ms * 1e3:8.1f} us  {2.0 * M * N * Kc / ms / 1e9:8.1f} TFLOP/s", flush=True)
        del Ac, Bc
for dil in (1, 2, 4, 8):
    for name, res, gn in (("conv1 gelu", None, False), ("conv2+GN", R16, True)):
        def run():
            rc = lib.tag_debug_gemm_tc(h, A.data_ptr(), K, Wt.data_ptr(), M, N, K, 5, dil, T, None, _lib.ptr(res), None, C16.data_ptr(), None, 1,
                                       _lib.ptr(gam if gn else None), _lib.ptr(bet if gn else None), None, None, s)
            _lib.check(h, rc, name)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"dil {dil} {name:12s}: {ms * 1e3:8.1f} us  {2.0 * M * N * K * 5 / ms / 1e9:8.1f} TFLOP/s", flush=True)
