"""What would the batch-1 ViT GEMMs cost with two 128-row M tiles (M=256) instead of three (M=257)?  Upper bound for a
'tail row on CUDA cores' variant.  Same method as gemm_graph_bench.py (200 launches in a CUDA graph)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv = [sys.argv[0]]
import gemm_graph_bench as gb  # runs its default table first (model plan only)

print("---- M=256 vs M=257, forced plans")
for (N, K, epi) in [(4224, 1408, 0), (6144, 1408, 1), (1408, 1408, 2), (1408, 6144, 2)]:
    for M in (257, 256):
        line = f"{M}x{N}x{K} epi{epi}:"
        os.environ.pop("VLFM_GEMM_FORCE", None)
        line += f" plan={gb.bench(M, N, K, epi):.2f}"
        for bn in (128, 64):
            for sp in ((1,) if epi != 2 else (2, 3, 4, 6, 8)):
                os.environ["VLFM_GEMM_FORCE"] = f"{bn}:{sp}"
                line += f" | {bn}:{sp}={gb.bench(M, N, K, epi):.2f}"
        print(line, flush=True)
