"""Randomized differential run of the CUDA path (functional API + ops, through the C-ABI) against the CPU oracle:
random ranks / shapes / channel axes / parameters / special values / alignments.  A development tool (the fixed cases
live in tests/); `tests/test_gpu_oracle.py::test_fuzz_vs_oracle` runs a short, seeded slice of it.

    python benchmarks/fuzz_vs_oracle.py [cases] [seed]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402  (the checker)
from qsparse_b200 import ops  # noqa: E402
from qsparse_b200._native import channel_layout  # noqa: E402
from qsparse_b200.quantize import quantize_with_decimal, quantize_with_line, quantize_with_scaler  # noqa: E402

SPECIALS = np.array([0.0, -0.0, 0.5, -0.5, 1.5, 2.5, -2.5, 1e-40, -1e-40, 3.96875, 4.0, -4.03125, 127.0, 128.0, -128.0,
                     0.3, -0.01, 100.0], np.float32)


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


def random_shape(rng):
    kind = rng.integers(0, 6)
    if kind == 0:      # NCHW feature maps, short rows
        hw = int(rng.choice([1, 3, 5, 7, 8, 13, 14, 16, 17, 28]))
        return (int(rng.integers(1, 9)), int(rng.integers(1, 70)), hw, int(rng.choice([hw, hw + 1, 1]))), 1
    if kind == 1:      # long rows
        return (int(rng.integers(1, 4)), int(rng.integers(1, 9)), int(rng.integers(200, 9000))), 1
    if kind == 2:      # [tokens, hidden], channel last
        return (int(rng.integers(1, 300)), int(rng.choice([1, 7, 8, 24, 64, 130, 1000]))), 1
    if kind == 3:      # weights, leading axis
        return (int(rng.integers(1, 40)), int(rng.integers(1, 20)), int(rng.choice([1, 3])), int(rng.choice([1, 3]))), 0
    if kind == 4:      # channel = last axis of a 3-D tensor
        return (int(rng.integers(1, 8)), int(rng.integers(1, 30)), int(rng.integers(1, 50))), 2
    return (int(rng.integers(1, 6)), int(rng.integers(1, 12)), int(rng.integers(1, 40)), int(rng.integers(1, 40))), \
        int(rng.integers(0, 4))


def random_input(rng, shape, scale, off):
    n = int(np.prod(shape))
    base = (rng.standard_normal(n + off) * scale).astype(np.float32)
    if n >= 4:
        idx = rng.choice(n, min(n, 12), replace=False) + off
        base[idx] = rng.choice(SPECIALS, idx.size)
    return base


def one_case(rng, case):
    shape, ci = random_shape(rng)
    off = int(rng.choice([0, 0, 1, 3]))                 # 4-byte-aligned views of the same data
    C = shape[ci]
    base = random_input(rng, shape, float(rng.choice([0.05, 1.0, 2.0, 30.0])), off)
    x = base[off:].reshape(shape)
    xc = cu(base)[off:].view(shape)
    layout = channel_layout(shape, ci)
    what = f"case {case} shape {shape} ci {ci} off {off}"

    dec = rng.integers(-2, 10, C).astype(np.float32)
    sc = rng.uniform(0.003, 0.9, C).astype(np.float32)
    lo = rng.uniform(-2, 0.2, C)
    lines = np.stack([lo, lo + rng.uniform(0.0, 3, C)], 1).astype(np.float32)
    if C > 2 and rng.random() < 0.3:
        lines[rng.integers(0, C)] = [0.25, 0.25]        # empty range -> step 1e-4
    bits = int(rng.choice([2, 4, 5, 8]))
    assert bits_equal(npy(quantize_with_decimal(xc, bits, cu(dec), ci)), orc.fq_pow2_fwd(x, dec, ci)), what
    assert bits_equal(npy(quantize_with_scaler(xc, bits, cu(sc), ci)), orc.fq_scaler_fwd(x, sc, ci)), what
    fzp = bool(rng.integers(0, 2))
    assert bits_equal(npy(quantize_with_line(xc, bits, cu(lines), ci, False, fzp)),
                      orc.fq_line_fwd(x, lines, bits, ci, fzp)), what + f" line fzp {fzp}"
    d0, s0 = float(rng.integers(0, 8)), float(rng.uniform(0.01, 0.5))
    assert bits_equal(npy(quantize_with_decimal(xc, bits, d0)), orc.fq_pow2_fwd(x, d0)), what
    assert bits_equal(npy(quantize_with_scaler(xc, bits, s0)), orc.fq_scaler_fwd(x, np.float32(s0))), what

    # channel-masked fused forward / backward
    cm = rng.random(C) > 0.5
    y = ops.fq_pow2_fwd(xc.contiguous(), cu(dec), layout, mask=cu(cm))
    assert bits_equal(npy(y), orc.fq_pow2_fwd(x, dec, ci, mask=cm)), what + " masked pow2"
    g = random_input(rng, shape, 3.0, 0).reshape(shape)
    g.reshape(-1)[::53] = np.nan
    for scale, is_dec in ((dec, True), (sc, False)):
        gc = cu(g).clone()
        notch = int(rng.integers(0, 2))
        ops.ste_bwd(gc, cu(scale), is_dec, bits, notch, layout)
        ref, _ = orc.ste_bwd(g, scale, bits, ci, is_dec, bool(notch))
        assert bits_equal(npy(gc), ref), what + f" ste is_dec {is_dec} notch {notch}"

    # statistics
    st = ops.reduce_stats(xc, layout, absmax=True, minmax=True, abssum=True, nnz=True)
    assert bits_equal(npy(st["absmax"]), orc.absmax(x, ci)), what + " absmax"
    mn, mx = orc.minmax(x, ci)
    assert np.array_equal(npy(st["min"]), mn) and np.array_equal(npy(st["max"]), mx), what + " minmax"
    o, c, i = layout
    xr = np.abs(x.astype(np.float64)).reshape(o, c, i)
    assert np.allclose(npy(st["abssum"]), xr.sum(axis=(0, 2)), rtol=3e-7, atol=0), what + " abssum"
    assert np.array_equal(npy(st["nnz"]), (xr != 0).sum(axis=(0, 2)).astype(np.float64)), what + " nnz"

    # exact select + unstructured mask
    n = x.size
    if n >= 2:
        k = int(rng.integers(0, n))
        flat = np.ascontiguousarray(x.reshape(-1))
        assert npy(ops.kth_value(xc.reshape(-1), k))[0] == np.sort(flat)[k], what + f" kth {k}"
        assert npy(ops.kth_value(xc.reshape(-1), k, take_abs=True))[0] == np.sort(np.abs(flat))[k], what + f" kth|.| {k}"
        sp = float(rng.choice([0.0, 0.25, 0.5, 0.9]))
        imp = np.abs(x)
        try:
            m_ref, thr_ref = orc.mask_given_importance(imp, sp)
        except IndexError:
            m_ref = None
        if m_ref is not None:
            mask = torch.empty(shape, dtype=torch.bool, device="cuda")
            thr = ops.kth_value(cu(imp), orc.kth_index(sp, n))
            yy = ops.mask_build_apply(cu(imp), thr, xc.contiguous(), mask)
            assert np.array_equal(npy(mask), m_ref), what + " mask"
            assert bits_equal(npy(yy), orc.mask_apply(x, m_ref.reshape(-1))), what + " mask apply"
    return what


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7fffffff), a)
    b = np.where(b < 0, -(b & 0x7fffffff), b)
    return np.abs(a - b)


def step_flow_case(rng, case):
    """the one-launch training step (statistics + parameter step in the last CTA) over a few steps of a random
    activation layout, against the oracle's squeeze / EMA / mask / scale / decimal flow"""
    kind = int(rng.integers(0, 3))
    if kind == 0:
        hw = int(rng.choice([7, 8, 14, 16, 28, 56]))
        outer, C, inner = int(rng.integers(1, 9)), int(rng.integers(2, 100)), hw * hw
    elif kind == 1:
        outer, C, inner = int(rng.integers(8, 600)), int(rng.choice([8, 24, 64, 130])), 1
    else:
        outer, C, inner = int(rng.integers(1, 5)), int(rng.integers(2, 40)), int(rng.integers(200, 5000))
    if outer * inner < 64:
        return None
    shape, layout = (outer, C, inner), (outer, C, inner)
    sparsity = float(rng.choice([0.25, 0.5, 0.75]))
    bits = int(rng.choice([4, 8]))
    k = orc.kth_index(sparsity, C)
    if k >= C:
        return None
    what = f"step flow {case} layout {layout} sparsity {sparsity} bits {bits}"
    gains = rng.uniform(0.05, 3.0, C).astype(np.float32).reshape(1, C, 1)
    mag = torch.zeros(C, device="cuda")
    mask = torch.ones(C, dtype=torch.bool, device="cuda")
    scale, dec = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    mag_ref, mask_ref, scale_ref = np.zeros(C, np.float32), np.ones(C, bool), np.zeros(1, np.float32)
    for t in range(3):
        x = np.maximum(rng.standard_normal(shape).astype(np.float32), 0) * gains
        xd = cu(x)
        ops.reduce_prune_quant_step(xd, layout, mag, mask, scale, dec, float(outer * inner), t, 1, t > 0, k, bits, t, True)
        mag_ref = orc.magnitude_ema(mag_ref, orc.squeeze_mean_abs(x, (1, C, 1)).reshape(-1), t)
        srt = np.sort(mag_ref)
        if np.min(np.diff(srt) / np.maximum(srt[1:], 1e-30)) < 1e-5:
            return None                                   # two magnitudes within the mean tolerance: the order is not defined
        if t > 0:
            mask_ref, _ = orc.mask_given_importance(mag_ref, sparsity)
        scale_ref = orc.scale_ema(scale_ref, np.array([np.max(orc.absmax(x, 1) * mask_ref)], np.float32), bits, t)
        assert np.array_equal(npy(mask), mask_ref), what + f" mask t {t}"
        assert bits_equal(npy(scale), scale_ref), what + f" scale t {t}"
        assert bits_equal(npy(dec), orc.scale_to_decimal(scale_ref)), what + f" decimal t {t}"
        assert ulp_diff(npy(mag), mag_ref).max() <= 8, what + f" magnitude t {t}"
        y = ops.fq_pow2_fwd(xd, dec, layout, mask=mask)
        assert bits_equal(npy(y), orc.fq_pow2_fwd(x, orc.scale_to_decimal(scale_ref), 1, mask=mask_ref)), what + f" y t {t}"
    return what


def row_quant_case(rng, case):
    """K8 (row-resident estimate + quantize, one launch) over a few EMA steps against the oracle"""
    rows = int(rng.integers(1, 400))
    inner = 8 * int(rng.choice([1, 2, 5, 16, 31, 32, 33, 100, 128, 129, 511, 512, 1025, 2048]))
    kind = str(rng.choice(["decimal", "scaler", "line"]))
    bits = int(rng.choice([2, 4, 8]))
    what = f"row quant {case} [{rows}, {inner}] {kind} bits {bits}"
    w = torch.zeros(rows, 2 if kind == "line" else 1, device="cuda")
    w_orc = np.zeros((rows, 2 if kind == "line" else 1), np.float32)
    for step in range(3):
        x = (rng.standard_normal((rows, inner)) * rng.choice([0.02, 1.0, 7.0])).astype(np.float32)
        if rows > 1 and step == 1:
            x[int(rng.integers(0, rows))] = 0.0
        xc = cu(x)
        if not ops.row_quant_supported(xc, rows):
            return None
        if kind == "line":
            t = step + 1
            y, _ = ops.row_quant_fused_(xc, w, ops.ROW_LINE, bits, t, True)
            mn, mx = orc.minmax(x, 0)
            w_orc = orc.lines_ema(w_orc, mn, mx, t)
            y_orc = orc.fq_line_fwd(x, w_orc, bits, 0, True)
        else:
            t = step
            y, d = ops.row_quant_fused_(xc, w, ops.ROW_DECIMAL if kind == "decimal" else ops.ROW_SCALER, bits, t)
            w_orc = orc.scale_ema(w_orc.reshape(-1), orc.absmax(x, 0), bits, t).reshape(rows, 1)
            if kind == "decimal":
                d_orc = orc.scale_to_decimal(w_orc.reshape(-1))
                assert bits_equal(npy(d), d_orc), what + f" decimal step {step}"
                y_orc = orc.fq_pow2_fwd(x, d_orc, 0)
            else:
                y_orc = orc.fq_scaler_fwd(x, w_orc.reshape(-1), 0)
        assert bits_equal(npy(w), w_orc), what + f" param step {step}"
        assert bits_equal(npy(y), y_orc), what + f" y step {step}"
    return what


def main(cases=300, seed=0):
    rng = np.random.default_rng(seed)
    for case in range(cases):
        one_case(rng, case)
    flows = rows_done = 0
    for case in range(max(cases // 4, 1)):
        flows += step_flow_case(rng, case) is not None
        rows_done += row_quant_case(rng, case) is not None
    print(f"fuzz: {flows} step flows and {rows_done} row-quant flows equal to the oracle")
    torch.cuda.synchronize()
    print(f"fuzz: {cases} random cases (seed {seed}) equal to the oracle")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
