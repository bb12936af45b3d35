"""One encoder-sized fused deformable-attention launch + the CCL kernels of one explore step, for `ncu --set full`."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vlfm_b200 import _lib
torch.manual_seed(0)
shapes = [(60, 80), (30, 40), (15, 20), (8, 10)]
b, heads, q = int(os.environ.get("B", "32")), 8, 6380
s = q
value = torch.randn(b, s, heads, 32, device="cuda").half()
offlog = torch.randn(b * q, 384, device="cuda")
ref = torch.rand(b, q, 4, 2, device="cuda")
out16 = torch.empty(b * q, 256, dtype=torch.float16, device="cuda")
arr = (ctypes.c_int32 * 8)(*[v for hw in shapes for v in hw])
lib = _lib.load()
def run():
    _lib.check(lib.vlfm_msda_fused(value.data_ptr(), offlog.data_ptr(), 384, 256, ref.data_ptr(), 2, out16.data_ptr(), b, s, q, heads, 4, 4,
                                   ctypes.cast(arr, ctypes.c_void_p), _lib.stream_ptr()), "msda")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
alg = value.numel() * 2 + offlog.numel() * 4 + ref.numel() * 4 + out16.numel() * 2
print(f"msda_fused B={b}: {ms*1e3:.1f} us/launch, algorithmic bytes {alg/1e6:.1f} MB -> {alg/ms/1e6:.0f} GB/s; gathered bytes {b*q*heads*16*4*64/1e9:.2f} GB -> {b*q*heads*16*4*64/ms/1e6:.0f} GB/s from L1/L2")
torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
