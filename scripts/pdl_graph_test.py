"""Does programmatic dependent launch survive torch CUDA-graph capture?  Times 400 dependent
LayerNorm launches (tiny kernels) as stream launches and as a graph, with VLFM_PDL=0/1 (env)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vlfm_b200 import _lib
lib = _lib.load()
x = torch.randn(257, 1408, device="cuda"); g = torch.ones(1408, device="cuda"); b = torch.zeros(1408, device="cuda")
o = torch.empty(257, 1408, device="cuda", dtype=torch.float16)
def ln():
    lib.vlfm_layernorm(x.data_ptr(), g.data_ptr(), b.data_ptr(), o.data_ptr(), None, 257, 1408, 1408, 1408, 0, 1e-6, _lib.stream_ptr())
def seq(n=400):
    for _ in range(n): ln()
def timeit(fn, it=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
print("PDL env", os.environ.get("VLFM_PDL", "1"))
print("stream: %.2f us/launch" % (timeit(seq) * 1e3 / 400))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    seq(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr): seq()
print("graph:  %.2f us/launch" % (timeit(gr.replay) * 1e3 / 400))
