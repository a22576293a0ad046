"""In-situ durations of the three kernels of the bench step (development tool; needs a library built with
QSB_EXTRA_NVCC_FLAGS=-DQSB_KERNEL_TIMING):  first-CTA-start / last-CTA-end %globaltimer stamps, no event records."""
import ctypes
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import _native as N  # noqa: E402
from qsparse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
lib = N.load_library()
SHAPE, LAYOUT, C = (256, 64, 56, 56), (256, 64, 3136), 64
gen = torch.Generator(device=dev).manual_seed(2)
x = torch.relu(torch.randn(SHAPE, device=dev, generator=gen))
g = torch.randn(SHAPE, device=dev, generator=gen)
y, gx = torch.empty_like(x), torch.empty_like(x)
st = dict(mag=torch.zeros(C, device=dev), mask=torch.ones(C, dtype=torch.bool, device=dev),
          scale=torch.zeros(1, device=dev), dec=torch.zeros(1, device=dev))
counter = torch.zeros(1, dtype=torch.int64, device=dev)
arrival = torch.zeros(64, dtype=torch.int32, device=dev)
k = 48
if len(sys.argv) > 1:
    ops.set_tuning(21, int(sys.argv[1]))
if len(sys.argv) > 2:
    ops.set_tuning(18, int(sys.argv[2]))


def step():
    ops.reduce_prune_quant_step(x, LAYOUT, st["mag"], st["mask"], st["scale"], st["dec"], 256 * 3136.0, 0, 1, 1, k, 8,
                                0, True, step_counter=counter, arrival=arrival)
    ops.fq_pow2_fwd(x, st["dec"], LAYOUT, mask=st["mask"], out=y)
    ops.ste_bwd(g, st["dec"], True, 8, 0, LAYOUT, mask=st["mask"], clamp_in_place=False, want_gx=True)


for _ in range(5):
    step()
side = torch.cuda.Stream(device=dev)
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
for _ in range(20):
    graph.replay()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (3 * 2 * 256))()
rows = []
for rep in range(30):
    for _ in range(3):
        graph.replay()
    N.check(lib.qsb_debug_kernel_times(None, N.stream_ptr(dev)), "reset")
    graph.replay()
    graph.replay()          # a successor, as in the timed loop
    torch.cuda.synchronize()
    N.check(lib.qsb_debug_kernel_times(ctypes.cast(buf, ctypes.c_void_p), None), "read")
    t = np.frombuffer(buf, dtype=np.uint64).reshape(3, 2, 256).astype(np.float64)
    s = [t[i, 0][t[i, 0] < 1.8e19].min() for i in range(3)]
    e = [t[i, 1].max() for i in range(3)]
    rows.append((s, e))
# the reset + two replays: stamps hold min-start of the FIRST replay and max-end of the SECOND; report per-kernel
# durations from a single-replay variant instead
rows = []
for rep in range(30):
    for _ in range(3):
        graph.replay()
    N.check(lib.qsb_debug_kernel_times(None, N.stream_ptr(dev)), "reset")
    graph.replay()
    torch.cuda.synchronize()
    N.check(lib.qsb_debug_kernel_times(ctypes.cast(buf, ctypes.c_void_p), None), "read")
    t = np.frombuffer(buf, dtype=np.uint64).reshape(3, 2, 256).astype(np.float64)
    s = [t[i, 0][t[i, 0] < 1.8e19].min() for i in range(3)]
    e = [t[i, 1].max() for i in range(3)]
    rows.append(s + e)
a = np.array(rows[5:]) / 1e3
s0, s1, s2, e0, e1, e2 = [a[:, i] for i in range(6)]
print(json.dumps({
    "args": sys.argv[1:],
    "what": "one graph replay of the bench step right after three others; us, mean of 25",
    "statistics_kernel": round(float((e0 - s0).mean()), 2), "gap_to_forward": round(float((s1 - e0).mean()), 2),
    "forward_kernel": round(float((e1 - s1).mean()), 2), "gap_to_backward": round(float((s2 - e1).mean()), 2),
    "backward_kernel": round(float((e2 - s2).mean()), 2), "first_start_to_last_end": round(float((e2 - s0).mean()), 2)}))
