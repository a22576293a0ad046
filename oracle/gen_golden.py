"""oracle/gen_golden.py — TEST INFRASTRUCTURE.

Generates ``tests/golden/*.npz`` by importing the UNMODIFIED reference
(mlzxy/qsparse v2.0.1) from /root/reference and running it on CPU tensors with
fixed seeds.  The reference cannot travel to the GPU box, the fixtures do.

    python oracle/gen_golden.py            # rewrites tests/golden/

Every array is small (the whole directory is < 2 MB).
"""
from __future__ import annotations

import io
import contextlib
import sys
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import qsparse  # noqa: F401  (the reference, not qsparse_b200)
    assert qsparse.__version__ == "2.0.1" and qsparse.__file__.startswith(REF)
    return qsparse


def adversarial(rng: np.random.Generator, shape, scale=1.0):
    """random normal with the awkward values of SURVEY §4 sprinkled in."""
    x = (rng.standard_normal(shape) * scale).astype(np.float32)
    flat = x.reshape(-1)
    specials = np.array(
        [0.0, -0.0, 1e-40, -1e-40, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 0.3, -0.01, 3.96875, 4.0, -4.0, -4.03125,
         127.0, 128.0, -128.0, 1e-8, 100.0, -100.0, 0.05, 0.15, 0.25, 0.35, -0.05, -0.15, 0.015625, 0.046875],
        dtype=np.float32)
    idx = rng.choice(flat.size, size=min(len(specials), flat.size), replace=False)
    flat[idx] = specials[: len(idx)]
    return x


def main():
    qsparse = import_reference()
    from qsparse.quantize import (quantize_with_decimal, quantize_with_scaler, quantize_with_line, DecimalQuantizer,
                                  ScalerQuantizer, AdaptiveQuantizer, DecimalQuantization, ScalerQuantization)
    from qsparse.sparse import MagnitudePruningCallback, PruneLayer
    from qsparse.util import squeeze_tensor_to_shape, calculate_mask_given_importance
    qsparse.set_qsparse_options(log_on_created=False)
    OUT.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(0)
    rng = np.random.default_rng(20261017)
    g = {}  # name -> array

    # ------------------------------------------------------------------ K1a pow2 forward
    x4 = adversarial(rng, (3, 5, 7, 6), 2.0)            # inner = 42 (not a multiple of 4)
    g["pow2/x"] = x4
    for d in (5, 0, -2, 12):
        g[f"pow2/y_d{d}"] = quantize_with_decimal(torch.from_numpy(x4), 8, d).numpy()
    dch = np.array([5, 3, 0, -1, 7], dtype=np.float32)
    g["pow2/dec_ch1"] = dch
    g["pow2/y_ch1"] = quantize_with_decimal(torch.from_numpy(x4), 8, torch.from_numpy(dch), 1).numpy()
    dch0 = np.array([4, 6, 2], dtype=np.float32)
    g["pow2/dec_ch0"] = dch0
    g["pow2/y_ch0"] = quantize_with_decimal(torch.from_numpy(x4), 8, torch.from_numpy(dch0), 0).numpy()
    # use_uint / flip_axis / bits have no effect on the forward (dead clamp)
    g["pow2/y_d5_uint_flip_b4"] = quantize_with_decimal(torch.from_numpy(x4), 4, 5, -1, True, False, True).numpy()
    x2 = adversarial(rng, (6, 10), 1.0)
    g["pow2/x2"] = x2
    d2 = rng.integers(-2, 9, size=10).astype(np.float32)
    g["pow2/dec2"] = d2
    g["pow2/y2_ch1"] = quantize_with_decimal(torch.from_numpy(x2), 8, torch.from_numpy(d2), 1).numpy()

    # ------------------------------------------------------------------ K1b scaler forward
    g["scaler/x"] = x4
    for name, s in (("0p1", 0.1), ("0p037", 0.037), ("0p5", 0.5), ("3", 3.0)):
        g[f"scaler/y_s{name}"] = quantize_with_scaler(torch.from_numpy(x4), 8, s).numpy()
    sch = np.array([0.1, 0.013, 0.5, 1.7, 0.0625], dtype=np.float32)
    g["scaler/s_ch1"] = sch
    g["scaler/y_ch1"] = quantize_with_scaler(torch.from_numpy(x4), 8, torch.from_numpy(sch), 1).numpy()

    # ------------------------------------------------------------------ K1c line forward
    g["line/x"] = x4
    for bits in (8, 4):
        for fzp in (True, False):
            g[f"line/y_tuple_b{bits}_fzp{int(fzp)}"] = quantize_with_line(
                torch.from_numpy(x4), bits, (-0.1, 0.9), -1, False, fzp).numpy()
    lines = np.stack([rng.uniform(-2, -0.1, 5), rng.uniform(0.1, 2, 5)], axis=1).astype(np.float32)
    lines[2] = (0.25, 0.25)                      # zero-width line -> step = 1e-4 (quantize.py:160)
    g["line/lines_ch1"] = lines
    for bits in (8, 4):
        for fzp in (True, False):
            g[f"line/y_ch1_b{bits}_fzp{int(fzp)}"] = quantize_with_line(
                torch.from_numpy(x4), bits, torch.from_numpy(lines), 1, False, fzp).numpy()

    # ------------------------------------------------------------------ K2 STE backward
    gr = adversarial(rng, (3, 5, 7, 6), 3.0)
    gr.reshape(-1)[[3, 77, 200]] = np.nan
    gr.reshape(-1)[[5, 99]] = [np.inf, -np.inf]
    g["bwd/g"] = gr

    def run_bwd(fn, *args):
        xin = torch.from_numpy(x4).clone().requires_grad_(True)
        y = fn(xin, *args)
        go = torch.from_numpy(gr).clone()
        y.backward(go)
        return xin.grad.numpy().copy(), go.numpy().copy()   # (grad wrt x, grad_output after the in-place clamp)

    for bits, d, flip in ((8, 5, False), (8, 5, True), (4, 3, False), (8, -1, False)):
        gx, go = run_bwd(quantize_with_decimal, bits, d, -1, False, False, flip)
        g[f"bwd/pow2_b{bits}_d{d}_f{int(flip)}_gx"] = gx
        g[f"bwd/pow2_b{bits}_d{d}_f{int(flip)}_go"] = go
    gx, go = run_bwd(quantize_with_decimal, 8, torch.from_numpy(dch), 1, False, False, False)
    g["bwd/pow2_ch1_gx"], g["bwd/pow2_ch1_go"] = gx, go
    gx, go = run_bwd(quantize_with_decimal, 8, 5, -1, False, True, False)   # backward_passthrough
    g["bwd/pow2_passthrough_gx"] = gx
    for name, s in (("0p1", 0.1), ("0p037", 0.037)):
        gx, go = run_bwd(quantize_with_scaler, 8, s, -1, False, False, False)
        g[f"bwd/scaler_s{name}_gx"] = gx
    gx, go = run_bwd(quantize_with_scaler, 6, torch.from_numpy(sch), 1, False, False, True)
    g["bwd/scaler_ch1_b6_flip_gx"] = gx

    # ------------------------------------------------------------------ K3/K4 DecimalQuantizer.optimize + decimal
    xs = [adversarial(rng, (2, 4, 6, 6), 0.5 + 0.7 * i) for i in range(4)]
    g["dq/xs"] = np.stack(xs)
    for bits in (8, 4):
        cb = DecimalQuantizer()
        w = torch.zeros(1, 1)
        ws = []
        for x in xs:                                  # per-tensor, activation (batched)
            nw = cb.optimize(torch.from_numpy(x), bits, w, batched=True, channel_index=-1)
            w.data[:] = nw
            ws.append(w.numpy().copy())
        g[f"dq/w_tensor_b{bits}"] = np.stack(ws)
    cb = ScalerQuantizer()
    w = torch.zeros(2, 1)
    ws = []
    for x in xs:                                      # weight path: channelwise=0, batched=False
        nw = cb.optimize(torch.from_numpy(x), 8, w, batched=False, channel_index=0)
        w.data[:] = nw
        ws.append(w.numpy().copy())
    g["dq/w_ch0"] = np.stack(ws)
    cb = ScalerQuantizer()
    w = torch.zeros(4, 1)
    ws = []
    for x in xs:                                      # weight path: channelwise=1 (transpose branch)
        nw = cb.optimize(torch.from_numpy(x), 8, w, batched=False, channel_index=1)
        w.data[:] = nw
        ws.append(w.numpy().copy())
    g["dq/w_ch1"] = np.stack(ws)
    # scale -> decimal  (quantize.py:316), dense around the sqrt(2) rounding boundaries
    base = np.float32(np.sqrt(0.5))
    near = np.array([np.nextafter(base, np.float32(0)), base, np.nextafter(base, np.float32(1))], dtype=np.float32)
    scales = np.concatenate([
        np.exp(rng.uniform(np.log(1e-6), np.log(1e3), 4000)).astype(np.float32),
        (near[None, :] * (2.0 ** np.arange(-12, 13))[:, None]).astype(np.float32).reshape(-1),
        (2.0 ** np.arange(-20, 21)).astype(np.float32),
        np.array([0.0, np.inf, -1.0, np.nan, 1e-45, 3e38], dtype=np.float32)])
    g["dq/scales"] = scales
    with np.errstate(all="ignore"):
        g["dq/decimals"] = (1 / torch.from_numpy(scales)).nan_to_num(posinf=1, neginf=1).log2().round().numpy()
    # full DecimalQuantizer.forward after optimize
    cb = DecimalQuantizer()
    w = torch.zeros(1, 1)
    x0 = torch.from_numpy(xs[0])
    w.data[:] = cb.optimize(x0, 8, w, batched=True, channel_index=-1)
    g["dq/fwd_tensor"] = cb(x0, 8, w, channel_index=-1).numpy()

    # ------------------------------------------------------------------ AdaptiveQuantizer.optimize
    for ci, batched, nch in ((1, True, 4), (0, False, 2), (-1, True, 1)):
        cb = AdaptiveQuantizer()
        w = torch.zeros(nch, 2)
        ws = []
        for x in xs:
            nw = cb.optimize(torch.from_numpy(x), 8, w, channel_index=ci, batched=batched)
            w.data[:] = nw
            ws.append(w.numpy().copy())
        g[f"aq/lines_ci{ci}_b{int(batched)}"] = np.stack(ws)

    # ------------------------------------------------------------------ squeeze_tensor_to_shape
    xq = adversarial(rng, (6, 8, 5, 7), 1.0)
    g["sq/x"] = xq
    for tgt in ((1, 8, 1, 1), (1, 8, 5, 7), (6, 1, 1, 1), (1, 8, 5, 1), (6, 8, 5, 7), (1, 1, 5, 7)):
        g["sq/" + "x".join(map(str, tgt))] = squeeze_tensor_to_shape(torch.from_numpy(xq).abs(), tgt).numpy()

    # ------------------------------------------------------------------ calculate_mask_given_importance
    imp = rng.random((10, 30, 7, 8)).astype(np.float32)
    g["mask/imp"] = imp
    for s in (0.0, 0.47, 0.5, 0.75, 0.999):
        g[f"mask/m_{s}"] = calculate_mask_given_importance(torch.from_numpy(imp), s).numpy()
    tie = np.array([0.5, 0.25, 0.25, 0.25, 0.75, 0.25, 1.0, 0.0, 0.0, 0.5, 0.5, 0.9], dtype=np.float32)
    g["mask/tie_imp"] = tie
    for s in (0.0, 0.25, 0.5, 0.75):
        g[f"mask/tie_m_{s}"] = calculate_mask_given_importance(torch.from_numpy(tie), s).numpy()
    neg = rng.standard_normal(257).astype(np.float32)
    g["mask/neg_imp"] = neg
    g["mask/neg_m_0.3"] = calculate_mask_given_importance(torch.from_numpy(neg), 0.3).numpy()

    # ------------------------------------------------------------------ MagnitudePruningCallback sequences
    def run_callback(shape, mask_shape, steps, sparsity, **kw):
        cb = MagnitudePruningCallback(**kw)
        cb.train()
        mask = torch.nn.Parameter(torch.ones(*mask_shape, dtype=torch.bool), requires_grad=False)
        xs_, outs, masks, mags = [], [], [], []
        for t in range(steps):
            x = torch.from_numpy(adversarial(rng, shape, 1.0 + 0.1 * t))
            xs_.append(x.numpy().copy())
            out = cb(x, sparsity, mask)
            outs.append(out.numpy().copy())
            masks.append(mask.data.numpy().copy())
            if hasattr(cb, "magnitude"):
                mags.append(cb.magnitude.data.numpy().copy())
        return np.stack(xs_), np.stack(outs), np.stack(masks), (np.stack(mags) if mags else np.zeros(0, np.float32))

    for tag, shape, mshape, kw in (
            ("struct", (4, 8, 5, 5), (1, 8, 1, 1), dict()),
            ("unstruct", (2, 3, 6, 6), (2, 3, 6, 6), dict()),
            ("struct_norunavg", (4, 8, 5, 5), (1, 8, 1, 1), dict(running_average=False)),
            ("struct_refresh2", (4, 8, 5, 5), (1, 8, 1, 1), dict(mask_refresh_interval=2, stop_mask_refresh=4))):
        xs_, outs, masks, mags = run_callback(shape, mshape, 6, 0.5, **kw)
        g[f"cb/{tag}_x"], g[f"cb/{tag}_out"], g[f"cb/{tag}_mask"], g[f"cb/{tag}_mag"] = xs_, outs, masks, mags

    # ------------------------------------------------------------------ PruneLayer ramp (docs/tutorial.ipynb:224-228)
    layer = PruneLayer(sparsity=0.5, start=200, interval=10, repetition=4, dimensions={1},
                       callback=MagnitudePruningCallback())
    layer.train()
    spars = []
    with contextlib.redirect_stdout(io.StringIO()):
        for step in range(241):
            layer(torch.from_numpy(rng.random((2, 8, 3, 3)).astype(np.float32)))
            spars.append(layer._cur_sparsity.item())
    g["ramp/cur_sparsity"] = np.array(spars, dtype=np.float64)
    g["ramp/final_mask"] = layer.mask.data.numpy().copy()

    # ------------------------------------------------------------------ layer-level flows
    from qsparse import quantize, prune
    data = torch.from_numpy(adversarial(rng, (1, 10, 12, 12), 1.0))
    g["layer/x"] = data.numpy()
    with contextlib.redirect_stdout(io.StringIO()):
        ql = quantize(bits=8, timeout=3, channelwise=-1, callback=DecimalQuantizer())
        outs = [ql(data).numpy().copy() for _ in range(6)]
        g["layer/q_dec_out"] = np.stack(outs)
        g["layer/q_dec_weight"] = ql.weight.data.numpy().copy()
        ql = quantize(bits=8, timeout=3, channelwise=1, callback=AdaptiveQuantizer())
        outs = [ql(data).numpy().copy() for _ in range(6)]
        g["layer/q_adp_out"] = np.stack(outs)
        g["layer/q_adp_weight"] = ql.weight.data.numpy().copy()
        ql.eval()
        g["layer/q_adp_eval_out"] = ql(data).numpy().copy()
        ql = quantize(bits=8, timeout=3, channelwise=-1)   # default ScalerQuantizer
        outs = [ql(data).numpy().copy() for _ in range(6)]
        g["layer/q_scl_out"] = np.stack(outs)
        g["layer/q_scl_weight"] = ql.weight.data.numpy().copy()
        pl = prune(sparsity=0.5, start=2, interval=2, repetition=3, dimensions={1})
        big = torch.from_numpy(adversarial(rng, (4, 10, 6, 6), 1.0))
        g["layer/px"] = big.numpy()
        outs, masks = [], []
        for _ in range(12):
            outs.append(pl(big).numpy().copy())
            masks.append(pl.mask.data.numpy().copy())
        g["layer/p_out"] = np.stack(outs)
        g["layer/p_mask"] = np.stack(masks)
        # weight path: quantize(prune(conv)) chain through imitate
        torch.manual_seed(7)
        conv = torch.nn.Conv2d(4, 6, 3)
        g["layer/conv_w"] = conv.weight.detach().numpy().copy()
        g["layer/conv_b"] = conv.bias.detach().numpy().copy()
        qp = quantize(prune(conv, sparsity=0.5, start=1, interval=1, repetition=2,
                            callback=MagnitudePruningCallback(running_average=False)),
                      bits=8, timeout=2, channelwise=0, callback=ScalerQuantizer())
        qp.train()
        ws = []
        for _ in range(6):
            ws.append(qp.weight.detach().numpy().copy())
        g["layer/qp_weight_seq"] = np.stack(ws)
        g["layer/qp_mask"] = qp.prune.mask.data.numpy().copy()
        g["layer/qp_scale"] = qp.quantize.weight.data.numpy().copy()

    np.savez_compressed(OUT / "hotpath_v1.npz", **g)
    total = sum(v.nbytes for v in g.values())
    print(f"wrote {len(g)} arrays, {total/1e6:.2f} MB raw -> {OUT/'hotpath_v1.npz'}")


if __name__ == "__main__":
    main()
