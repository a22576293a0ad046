"""One-off timing of the STOCK reference (mlzxy/qsparse, imported from /root/reference, unmodified) on the
build container's CPU for BASELINE config 2: Sequential(PruneLayer(0.75, dimensions={1}), QuantizeLayer(8,
channelwise=-1, DecimalQuantizer())) forward + backward on [256,64,56,56] in the steady state.  Runs only where
/root/reference exists; the result is committed as profiles/r02_reference_cpu_timing.json.

    python benchmarks/time_reference_cpu.py > profiles/r02_reference_cpu_timing.json
"""
import contextlib
import io
import json
import os
import sys
import time

import torch

sys.path.insert(0, "/root/reference")
with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
    import qsparse
    from qsparse.quantize import DecimalQuantizer
    from qsparse.sparse import MagnitudePruningCallback
assert qsparse.__file__.startswith("/root/reference")
qsparse.set_qsparse_options(log_on_created=False)
threads = len(os.sched_getaffinity(0))
torch.set_num_threads(threads)
torch.manual_seed(2)
shape = (256, 64, 56, 56)
x = torch.relu(torch.randn(shape))
g = torch.randn(shape)
with contextlib.redirect_stdout(io.StringIO()):
    p = qsparse.prune(sparsity=0.75, dimensions={1}, start=0, interval=1, repetition=1, callback=MagnitudePruningCallback())
    q = qsparse.quantize(bits=8, channelwise=-1, timeout=1, callback=DecimalQuantizer())
    p.train(), q.train()

    def step():
        xr = x.clone().requires_grad_(True)
        t0 = time.perf_counter()
        y = q(p(xr))
        t1 = time.perf_counter()
        y.backward(g.clone())
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    for _ in range(3):
        step()
    f, b = [], []
    for _ in range(5):
        a, c = step()
        f.append(a), b.append(c)
n = x.numel()
fwd, bwd = min(f), min(b)
print(json.dumps({
    "what": "the UNMODIFIED reference (import qsparse from /root/reference) on CPU tensors: q(p(x)) forward + backward, "
            "steady state (pruning active, quantizer past its timeout), best of 5",
    "shape": list(shape), "threads": threads, "torch": torch.__version__,
    "forward_ms": round(fwd * 1e3, 1), "backward_ms": round(bwd * 1e3, 1), "step_ms": round((fwd + bwd) * 1e3, 1),
    "gbs_on_20_bytes_per_elem": round(20 * n / (fwd + bwd) / 1e9, 2),
    "note": "the plain-C port timed by bench.py's cpu_baseline / --impl reference does the same step at ~21-35 GB/s on "
            "16 threads of the GPU box: the port is the FASTER CPU baseline, the GPU/CPU ratios quoted against it are "
            "conservative"}))
