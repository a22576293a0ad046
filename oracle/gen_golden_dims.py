"""oracle/gen_golden_dims.py — TEST INFRASTRUCTURE.

Prune masks whose kept axes are NOT adjacent (``dimensions={1, 3}`` on NCHW activations, ``{0, 2}`` on conv
weights, ``{0, 3}`` ...): the reference takes any dimension set through broadcasting (qsparse/sparse.py:66,
util.py:93-101).  Runs the UNMODIFIED reference from /root/reference on CPU with fixed seeds and records, per case
and step, the layer output, the mask, the running magnitude and the input gradient.

    python oracle/gen_golden_dims.py        # rewrites tests/golden/dims_v1.npz  (build container only)
"""
from __future__ import annotations

import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from gen_golden import import_reference, OUT  # noqa: E402

CASES = {
    # name: (input shape, dimensions, running_average)
    "nchw_13": ((4, 6, 5, 8), {1, 3}, True),
    "nchw_03": ((4, 6, 5, 8), {0, 3}, True),
    "w_02": ((8, 4, 3, 3), {0, 2}, True),
    "w_02_instant": ((8, 4, 3, 3), {0, 2}, False),
    "w_013": ((8, 4, 3, 5), {0, 1, 3}, True),
    "nlc_02": ((6, 7, 16), {0, 2}, True),
}
STEPS = 8


def main():
    import_reference()
    from qsparse import prune
    from qsparse.sparse import MagnitudePruningCallback
    g = {}
    for ci, (name, (shape, dims, ra)) in enumerate(CASES.items()):
        rng = np.random.default_rng(100 + ci)
        xs = rng.standard_normal((STEPS,) + shape).astype(np.float32)
        gs = rng.standard_normal((STEPS,) + shape).astype(np.float32)
        g[f"{name}/x"], g[f"{name}/g"] = xs, gs
        with contextlib.redirect_stdout(io.StringIO()):
            pl = prune(sparsity=0.5, start=2, interval=2, repetition=2, dimensions=dims,
                       callback=MagnitudePruningCallback(running_average=ra))
            pl.train()
            outs, masks, mags, gxs = [], [], [], []
            for t in range(STEPS):
                x = torch.from_numpy(xs[t]).requires_grad_(True)
                y = pl(x)
                y.backward(torch.from_numpy(gs[t]))
                outs.append(y.detach().numpy().copy())
                gxs.append(x.grad.numpy().copy())
                masks.append(pl.mask.data.numpy().copy())
                if ra:
                    mags.append(pl.callback.magnitude.data.numpy().copy() if hasattr(pl.callback, "magnitude")
                                else np.zeros(pl.mask.shape, np.float32))
        g[f"{name}/out"], g[f"{name}/gx"], g[f"{name}/mask"] = np.stack(outs), np.stack(gxs), np.stack(masks)
        if ra:
            g[f"{name}/mag"] = np.stack(mags)
    np.savez_compressed(OUT / "dims_v1.npz", **g)
    print("wrote", OUT / "dims_v1.npz", len(g), "arrays", sum(v.nbytes for v in g.values()) / 1e3, "KB raw")


if __name__ == "__main__":
    main()
