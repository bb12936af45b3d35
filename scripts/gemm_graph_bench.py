# NOTE: needs a development build of the library (NVCC flag -DVLFM_DEV_PROBES): the probe entry points are not in the shipped C-ABI.
"""Per-shape GEMM timing at full clocks: 200 back-to-back launches captured in a CUDA graph
(no CPU launch bound), optional sweep of the tile plan via VLFM_GEMM_FORCE=bn:splits."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vlfm_b200.vlm.dense import gemm_f16

SHAPES = [(257, 4224, 1408, 0), (257, 1408, 1408, 2), (257, 6144, 1408, 1), (257, 1408, 6144, 2),
          (32, 2304, 768, 0), (32, 768, 768, 2), (32, 3072, 768, 1), (32, 768, 3072, 2), (257, 9216, 1408, 0)]
N = 200

def bench(M, Nn, K, epi):
    # distinct weights per launch (ring of 8) so that weights stream from HBM like in the real forward
    a = torch.randn(M, K, device="cuda").half()
    ws = [torch.randn(Nn, K, device="cuda").half() for _ in range(8)]
    b = torch.zeros(Nn, device="cuda")
    o = torch.zeros(M, Nn, device="cuda", dtype=torch.float32 if epi >= 2 else torch.float16)
    def seq():
        for i in range(N): gemm_f16(a, ws[i % 8], b, epi, o)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        seq(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): seq()
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * N)

sweep = len(sys.argv) > 1 and sys.argv[1] == "sweep"
for (M, Nn, K, epi) in SHAPES:
    os.environ.pop("VLFM_GEMM_FORCE", None)
    base = bench(M, Nn, K, epi)
    line = f"{M}x{Nn}x{K} epi{epi}: model-plan {base:6.2f} us ({2*M*Nn*K/base/1e6:6.1f} TF)"
    if sweep:
        for bn in (128, 64, 32):
            for sp in ((1, 2, 3, 4, 6, 8) if epi == 2 else (1,)):
                os.environ["VLFM_GEMM_FORCE"] = f"{bn}:{sp}"
                line += f" | {bn}:{sp}={bench(M, Nn, K, epi):.2f}"
                if len(sys.argv) > 2 and sp <= 4:
                    os.environ["VLFM_GEMM_FORCE"] = f"{bn}:{sp}:1"
                    line += f" {bn}:{sp}:sh={bench(M, Nn, K, epi):.2f}"
    print(line, flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "timeline":
    # phases of CTA (0,0,0) of the LAST launch of a 200-launch graph (sustained clocks)
    import numpy as np
    from vlfm_b200 import _lib
    lib = _lib.load()
    buf = torch.zeros(8 + 4096, dtype=torch.int64, device="cuda")
    names = ["start", "setup", "depwait", "stage0(w1)", "lastmma(w1)", "accready(w2)", "epidone"]
    for (M, Nn, K, epi) in SHAPES[:4]:
        os.environ.pop("VLFM_GEMM_FORCE", None)
        lib.vlfm_gemm_debug_timeline(buf.data_ptr())
        t_us = bench(M, Nn, K, epi)
        lib.vlfm_gemm_debug_timeline(None)
        t = buf.cpu().tolist()
        st = np.array(t[8:8 + 4096:2]); en = np.array(t[9:9 + 4096:2]); ok = st > 0
        rel = [(t[i] - t[0]) / 1.965e3 for i in range(7)]
        print(f"{M}x{Nn}x{K} epi{epi}: {t_us:.2f} us/launch | " + " ".join(f"{n}={v:.2f}" for n, v in zip(names, rel)) +
              f" | ctas={ok.sum()} skew={(st[ok].max()-st[ok].min())/1e3:.2f} life_avg={(en[ok]-st[ok]).mean()/1e3:.2f} life_max={(en[ok]-st[ok]).max()/1e3:.2f} span={(en[ok].max()-st[ok].min())/1e3:.2f}")
        buf.zero_()
