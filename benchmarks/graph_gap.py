"""How much of the fused step is launch gaps?  Times the 5-launch step eagerly and as a
captured CUDA graph (same kernels, same arguments)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(2)
shape, layout = (256, 64, 56, 56), (256, 64, 3136)
x = torch.relu(torch.randn(shape, device=dev))
g = torch.randn(shape, device=dev)
y = torch.empty_like(x)
gx = torch.empty_like(x)
mag = torch.zeros(64, device=dev)
mask = torch.ones(64, dtype=torch.bool, device=dev)
scale = torch.zeros(1, device=dev)
dec = torch.zeros(1, device=dev)


def step(t=5):
    st = ops.reduce_stats(x, layout, abssum=True, absmax=True)
    ops.prune_quant_params(mag, mask, scale, dec, st, 256 * 3136.0, t, 1, t > 0, 48, 8, t, True)
    ops.fq_pow2_fwd(x, dec, layout, mask=mask, out=y)
    ops.ste_bwd(g, dec, True, 8, 0, layout, mask=mask, clamp_in_place=False, want_gx=True)


def timeit(fn, n=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / n * 1e3


for t in range(3):
    step(t)
print(f"eager : {timeit(step):8.1f} us/step")
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
print(f"graph : {timeit(graph.replay):8.1f} us/step")
