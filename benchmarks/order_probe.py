"""Does walking a tensor BACKWARDS in the map kernel that follows a front-to-back reduction of the same tensor turn
DRAM reads into L2 hits on B200?  (tuning key 2.)  Times the pair {per-channel abs-max statistics -> pow2 fake-quant
forward} as one CUDA graph, many tensor sizes, natural and reversed tile order.  Development probe."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
C, INNER = 64, 3136
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
rows = []
for outer in (32, 64, 96, 128, 160, 192, 256, 384, 512):
    layout = (outer, C, INNER)
    x = torch.randn(outer, C, INNER, device=dev)
    y = torch.empty_like(x)
    dec = torch.full((1,), 5.0, device=dev)
    res = {}
    for rev in (0, 1):
        ops.set_tuning(2, rev)
        out = ops.reduce_stats(x, layout, absmax=True)

        def pair():
            ops.reduce_stats(x, layout, absmax=True, out=out)
            ops.fq_pow2_fwd(x, dec, layout, out=y)

        for _ in range(3):
            pair()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            pair()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            pair()
        ts = []
        for _ in range(30):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        res[rev] = ts[len(ts) // 2]
    rows.append({"mb": x.numel() * 4 / 2**20, "natural_us": round(res[0], 2), "reversed_us": round(res[1], 2)})
    print(rows[-1], flush=True)
    del x, y
print(json.dumps({"what": "abs-max statistics -> pow2 forward on the same tensor, one graph, L2 flushed before", "rows": rows}))
