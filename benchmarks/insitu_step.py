"""The bench step, eagerly, a fixed number of times: the target of
  ncu --replay-mode application --cache-control none --clock-control none -k regex:'reduce_rows|map_chan' -s 30 -c 9 ...
which reports each kernel's DRAM bytes and L2 hit rate with the cache state its predecessor left behind (kernel
replay flushes or at least re-runs the kernel on its own leftovers).  Development tool."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
SHAPE, LAYOUT, C = (256, 64, 56, 56), (256, 64, 3136), 64
gen = torch.Generator(device=dev).manual_seed(2)
x = torch.relu(torch.randn(SHAPE, device=dev, generator=gen))
g = torch.randn(SHAPE, device=dev, generator=gen)
y, gx = torch.empty_like(x), torch.empty_like(x)
st = dict(mag=torch.zeros(C, device=dev), mask=torch.ones(C, dtype=torch.bool, device=dev),
          scale=torch.zeros(1, device=dev), dec=torch.zeros(1, device=dev))
counter = torch.zeros(1, dtype=torch.int64, device=dev)
arrival = torch.zeros(64, dtype=torch.int32, device=dev)
if len(sys.argv) > 1:
    ops.set_tuning(18, int(sys.argv[1]))
if len(sys.argv) > 2:
    ops.set_tuning(2, int(sys.argv[2]))
torch.cuda.synchronize()
for _ in range(14):
    ops.reduce_prune_quant_step(x, LAYOUT, st["mag"], st["mask"], st["scale"], st["dec"], 256 * 3136.0, 0, 1, 1, 48, 8,
                                0, True, step_counter=counter, arrival=arrival)
    ops.fq_pow2_fwd(x, st["dec"], LAYOUT, mask=st["mask"], out=y)
    ops.ste_bwd(g, st["dec"], True, 8, 0, LAYOUT, mask=st["mask"], clamp_in_place=False, want_gx=True)
torch.cuda.synchronize()
print("done", float(st["dec"]))
