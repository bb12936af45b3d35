"""Quick on-GPU micro-timings (CUDA events) for development; not the bench."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vlfm_b200 import _lib
from vlfm_b200.vlm.dense import gemm_f16


def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    print(torch.cuda.get_device_name(0))
    for (M, N, K) in [(257, 4224, 1408), (257, 1408, 1408), (257, 6144, 1408), (257, 1408, 6144),
                      (8224, 4224, 1408), (8224, 1408, 1408), (8224, 6144, 1408), (8224, 1408, 6144), (8192, 8192, 8192)]:
        a = torch.randn(M, K, device="cuda").half(); w = torch.randn(N, K, device="cuda").half(); b = torch.zeros(N, device="cuda")
        o = torch.empty(M, N, device="cuda", dtype=torch.float16)
        ms = timeit(lambda: gemm_f16(a, w, b, 0, o))
        ms_t = timeit(lambda: torch.matmul(a, w.t()))
        print(f"gemm {M}x{N}x{K}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s   (torch/cuBLAS {ms_t*1e3:8.1f} us {2*M*N*K/ms_t/1e9:8.1f} TF)")
    for (M, N, K) in [(257, 1408, 1408), (257, 1408, 6144), (32, 768, 3072)]:
        a = torch.randn(M, K, device="cuda").half(); w = torch.randn(N, K, device="cuda").half(); b = torch.zeros(N, device="cuda")
        o = torch.zeros(M, N, device="cuda", dtype=torch.float32)
        ms = timeit(lambda: gemm_f16(a, w, b, 2, o))
        print(f"gemm+resid(split-K) {M}x{N}x{K}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s")
    from vlfm_b200.mapping.value_map import ValueMapBatch
    from vlfm_b200.utils.synthetic import trajectory
    fov = float(np.deg2rad(79))
    for B in (1, 32):
        eng = ValueMapBatch(B, 1, size=1000, use_max_confidence=False)
        fr = trajectory(1, 1)[0]
        depth = torch.from_numpy(np.stack([fr.depth] * B)).cuda(); tf = torch.from_numpy(np.stack([fr.tf] * B)).cuda()
        vals = torch.full((B, 1), 0.5, dtype=torch.float64).cuda()
        ms = timeit(lambda: eng.update(vals, depth, tf, 0.5, 5.0, fov))
        print(f"value update B={B}: {ms*1e3:.1f} us/step-batch, {B/ms*1e3:.0f} env-steps/s, alg bytes {2.04e6*B/ms/1e6:.1f} GB/s")


if __name__ == "__main__":
    main()
