# NOTE: needs a development build of the library (NVCC flag -DVLFM_DEV_PROBES): the probe entry points are not in the shipped C-ABI.
"""Per-phase timeline of CTA (0,0,0) of the tcgen05 GEMM (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vlfm_b200 import _lib
from vlfm_b200.vlm.dense import gemm_f16
lib = _lib.load()
buf = torch.zeros(8 + 4096, dtype=torch.int64, device="cuda")
names = ["start", "setup", "depwait", "stage0", "lastmma", "accready", "epidone"]
for (M, N, K, epi) in [(257, 4224, 1408, 0), (257, 1408, 1408, 2), (257, 6144, 1408, 1), (257, 1408, 6144, 2), (8224, 6144, 1408, 1), (32, 768, 768, 0)]:
    a = torch.randn(M, K, device="cuda").half(); w = torch.randn(N, K, device="cuda").half(); b = torch.zeros(N, device="cuda")
    o = torch.zeros(M, N, device="cuda", dtype=torch.float32 if epi >= 2 else torch.float16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for rep in range(3):
        flush.zero_(); torch.cuda.synchronize()
        lib.vlfm_gemm_debug_timeline(buf.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gemm_f16(a, w, b, epi, o); e1.record(); torch.cuda.synchronize()
        lib.vlfm_gemm_debug_timeline(None)
        t = buf.cpu().tolist()
        rel = [(t[i] - t[0]) / 1.965e3 for i in range(7)]
        import numpy as np
        nct = ((N + 127) // 128) * ((M + 127) // 128)
        st = np.array(t[8:8 + 2 * 2048:2]); en = np.array(t[9:9 + 2 * 2048:2]); ok = st > 0
        st, en = st[ok], en[ok]; buf.zero_()
        skew = f"ctas={ok.sum()} start-skew={(st.max()-st.min())/1e3:.2f}us life(avg)={(en-st).mean()/1e3:.2f}us span={(en.max()-st.min())/1e3:.2f}us"
    print(f"{M}x{N}x{K} epi{epi}: event {e0.elapsed_time(e1)*1e3:.1f} us | " + " ".join(f"{n}={v:.2f}" for n, v in zip(names, rel)) + " | " + skew)
