"""oracle/gen_golden_extremes.py — TEST INFRASTRUCTURE.

Golden vectors for the corners the reference's own tests never visit, produced by the UNMODIFIED reference
(mlzxy/qsparse v2.0.1, imported from /root/reference) on CPU tensors:

  * decimals / scales / line ranges far outside the useful range (2^d overflowing, zero / negative / infinite /
    NaN scales, empty and inverted ranges) over inputs with +-0, +-inf, NaN, subnormals and huge values;
  * bit widths 1 ... 32 for the line quantizer and the straight-through backward (2^bits - 1 is not
    representable in fp32 from 25 bits on; at 1 bit a clamp bound is +0.0 and the sign of a zero gradient shows);
  * calculate_mask_given_importance on NaN / inf / constant / one- and two-element importances at sparsity 0 ... 1.

    python oracle/gen_golden_extremes.py        # rewrites tests/golden/extremes_v1.npz  (< 1 MB)
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from gen_golden import import_reference, OUT  # noqa: E402

DECS = [-40.0, -1.0, 0.0, 30.0, 100.0, 126.0, 130.0]
SCALES = [1e-30, 3e37, -0.5, 0.0, float("inf"), float("nan"), 1e-42]
LINES = [(0.0, 0.0), (1.0, 1.0), (0.5, -0.5), (0.0, 1e-30), (-1e30, 1e30), (0.0, float("inf")), (-0.3, 0.7)]
BITS = [1, 2, 3, 12, 16, 24, 32]
MASK_SPARSITIES = [0.0, 0.3, 0.5, 0.75, 0.999, 1.0]
SPECIAL = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e38, -1e38, 1e-40, -1e-40, 3e9, -3e9, 0.5], np.float32)


def main():
    qsparse = import_reference()
    from qsparse.quantize import quantize_with_decimal, quantize_with_scaler, quantize_with_line
    qsparse.set_qsparse_options(log_on_created=False)
    rng = np.random.default_rng(20261018)
    g = {}
    C = len(DECS)
    x = (rng.standard_normal((3, C, 330)) * 2).astype(np.float32)
    x[:, :, :12] = SPECIAL
    x[:, :, 12:40] *= np.float32(1e-12)
    g["x"] = x
    xt = torch.from_numpy(x)
    g["decs"], g["scales"], g["lines"] = (np.array(DECS, np.float32), np.array(SCALES, np.float32),
                                          np.array(LINES, np.float32))
    g["pow2/ch"] = quantize_with_decimal(xt.clone(), 8, torch.tensor(DECS), 1).numpy()
    g["scaler/ch"] = quantize_with_scaler(xt.clone(), 8, torch.tensor(SCALES), 1).numpy()
    for i, d in enumerate(DECS):
        g[f"pow2/t{i}"] = quantize_with_decimal(xt.clone(), 8, d).numpy()
    for i, s in enumerate(SCALES):
        g[f"scaler/t{i}"] = quantize_with_scaler(xt.clone(), 8, s).numpy()
    for fzp in (True, False):
        g[f"line/ch_fzp{int(fzp)}"] = quantize_with_line(xt.clone(), 8, torch.tensor(LINES), 1, False, fzp).numpy()
        for i, ln in enumerate(LINES):
            g[f"line/t{i}_fzp{int(fzp)}"] = quantize_with_line(xt.clone(), 8, ln, -1, False, fzp).numpy()

    # bit widths
    xb = (rng.standard_normal((3, 5, 41)) * 2).astype(np.float32)
    xb.reshape(-1)[::53] *= 1e6
    xb.reshape(-1)[[1, 2, 3, 4]] = [0.0, -0.0, 0.5, -0.5]
    gb = (rng.standard_normal((3, 5, 41)) * 3).astype(np.float32)
    gb.reshape(-1)[::31] *= 1e9
    gb.reshape(-1)[[1, 2, 3, 4, 5]] = [0.0, -0.0, np.nan, np.inf, -np.inf]
    g["bits/x"], g["bits/g"] = xb, gb
    lines5 = np.stack([rng.uniform(-2, -0.1, 5), rng.uniform(0.1, 2, 5)], 1).astype(np.float32)
    dec5 = rng.integers(-2, 7, 5).astype(np.float32)
    sc5 = rng.uniform(0.001, 0.05, 5).astype(np.float32)
    g["bits/lines"], g["bits/dec"], g["bits/scale"] = lines5, dec5, sc5
    for b in BITS:
        for fzp in (True, False):
            g[f"bits/line_b{b}_fzp{int(fzp)}"] = quantize_with_line(
                torch.from_numpy(xb), b, torch.from_numpy(lines5), 1, False, fzp).numpy()
        for name, fn, par in (("dec", quantize_with_decimal, dec5), ("scale", quantize_with_scaler, sc5)):
            for flip in (False, True):
                xin = torch.from_numpy(xb).clone().requires_grad_(True)
                go = torch.from_numpy(gb).clone()
                fn(xin, b, torch.from_numpy(par), 1, False, False, flip).backward(go)
                g[f"bits/bwd_{name}_b{b}_f{int(flip)}"] = go.numpy().copy()       # grad_output after the in-place clamp
    # masks from importances with NaN / inf / constant / tiny tensors at the ends of the sparsity range
    from qsparse.util import calculate_mask_given_importance
    base = rng.standard_normal(1000).astype(np.float32)
    nan7 = base.copy()
    nan7[::7] = np.nan
    infs = base.copy()
    infs[:5] = [np.inf, -np.inf, np.inf, 0.0, -0.0]
    mask_cases = {"normal": base, "nan": nan7, "inf": infs, "allnan": np.full(64, np.nan, np.float32),
                  "const": np.full(100, 0.25, np.float32), "two": np.array([3.0, 1.0], np.float32),
                  "one": np.array([3.0], np.float32)}
    g["mask/names"] = np.array(list(mask_cases))
    g["mask/sparsities"] = np.array(MASK_SPARSITIES)
    for name, v in mask_cases.items():
        g[f"mask/{name}/imp"] = v
        for i, sp in enumerate(MASK_SPARSITIES):
            try:
                g[f"mask/{name}/s{i}"] = calculate_mask_given_importance(torch.from_numpy(v), sp).numpy()
            except IndexError:
                g[f"mask/{name}/s{i}"] = np.array([-1], np.int8)          # the reference raises IndexError
    # scale -> decimal on degenerate scales: round(log2(nan_to_num(1/s, posinf=1, neginf=1)))  (quantize.py:316)
    sd = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 1e-40, -0.5, -1e-30, 3e38, 1.0, 0.7071068, 1.4142135,
                   2.0 ** -126, 2.0 ** 127, 2.0 ** -149, 1.1754942e-38, 0.1, 3.0], np.float32)
    g["s2d/scales"] = sd
    g["s2d/decimals"] = (1 / torch.from_numpy(sd.copy())).nan_to_num(posinf=1, neginf=1).log2().round().numpy()
    np.savez_compressed(OUT / "extremes_v1.npz", **g)
    print("wrote", OUT / "extremes_v1.npz", len(g), "arrays")


if __name__ == "__main__":
    main()
