"""Timeline of one CTA pair of the fused transformer-layer tail (experiments build): clock64 stamps of the MMA issuer, the first
epilogue warp and the TMA producer of CTA 0. Run: python tools/run_exp.py tools/tl_trace.py"""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tag_b200 as tb
from tag_b200 import _lib

DEV = "cuda:0"
lib = _lib.load()
h = tb.scoring.util_handle(DEV)
M, F, D = 412500, 1024, 256
g = torch.Generator(device=DEV).manual_seed(1)
att = torch.randn(M, D, device=DEV, generator=g).half()
X = torch.randn(M, D, device=DEV, generator=g)
Wo = (torch.randn(D, D, device=DEV, generator=g) / 16).half()
W1 = (torch.randn(F, D, device=DEV, generator=g) / 16).half()
W2 = (torch.randn(D, F, device=DEV, generator=g) / 32).half()
bo, b1, b2 = (0.1 * torch.randn(n, device=DEV, generator=g) for n in (D, F, D))
g1, be1, g2, be2 = (torch.ones(D, device=DEV), torch.zeros(D, device=DEV), torch.ones(D, device=DEV), torch.zeros(D, device=DEV))
X16 = torch.empty(M, D, device=DEV, dtype=torch.float16)
s = torch.cuda.current_stream().cuda_stream

def run():
    rc = lib.tag_debug_tlayer_tail(h, att.data_ptr(), X.data_ptr(), X16.data_ptr(), M, F, Wo.data_ptr(), W1.data_ptr(), W2.data_ptr(), bo.data_ptr(),
                                   b1.data_ptr(), b2.data_ptr(), g1.data_ptr(), be1.data_ptr(), g2.data_ptr(), be2.data_ptr(), s)
    _lib.check(h, rc, "tail")

for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print("tail kernel: %.1f us per launch, %.0f TFLOP/s" % (e0.elapsed_time(e1) * 100, (2.0 * M * D * D + 4.0 * M * F * D) / (e0.elapsed_time(e1) / 10) / 1e9))
buf = torch.zeros(3 * 4096, dtype=torch.int64, device=DEV)
fn = lib.tag_exp_set_tlayer_trace
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.data_ptr()) == 0
run()
torch.cuda.synchronize()
fn(None)
t = buf.cpu().view(3, 2048, 2)
names = {0: "mma", 1: "epi", 2: "tma"}
ev = []
for role in range(3):
    for tag, clk in t[role].tolist():
        if clk:
            ev.append((clk, names[role], tag))
ev.sort()
t0 = ev[0][0]
# print tiles 2 and 3 of the pair (steady state): find the 3rd occurrence of (mma, 1)
starts = [i for i, e in enumerate(ev) if e[1] == "mma" and e[2] == 1]
lo, hi = starts[2], starts[4]
only = os.environ.get("TL_ROLE")          # e.g. TL_ROLE=epi: one role only
for clk, who, tag in ev[lo - 12:hi]:
    if only is None or who == only:
        print(f"{clk - ev[lo][0]:8d}  {who}  {tag}")
