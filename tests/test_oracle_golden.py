"""Pins the CPU oracle (oracle/) against golden vectors produced by the imported
reference (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.conftest import bits_equal, ulp_diff


def test_pow2_forward(golden):
    g = golden
    x = g["pow2/x"]
    for d in (5, 0, -2, 12):
        assert bits_equal(orc.fq_pow2_fwd(x, d), g[f"pow2/y_d{d}"]), d
    assert bits_equal(orc.fq_pow2_fwd(x, g["pow2/dec_ch1"], 1), g["pow2/y_ch1"])
    assert bits_equal(orc.fq_pow2_fwd(x, g["pow2/dec_ch0"], 0), g["pow2/y_ch0"])
    # bits / use_uint / flip_axis do not change the forward (dead clamp, quantize.py:56-62)
    assert bits_equal(orc.fq_pow2_fwd(x, 5), g["pow2/y_d5_uint_flip_b4"])
    assert bits_equal(orc.fq_pow2_fwd(g["pow2/x2"], g["pow2/dec2"], 1), g["pow2/y2_ch1"])


def test_scaler_forward(golden):
    g = golden
    x = g["scaler/x"]
    for name, s in (("0p1", 0.1), ("0p037", 0.037), ("0p5", 0.5), ("3", 3.0)):
        assert bits_equal(orc.fq_scaler_fwd(x, np.float32(s)), g[f"scaler/y_s{name}"]), name
    assert bits_equal(orc.fq_scaler_fwd(x, g["scaler/s_ch1"], 1), g["scaler/y_ch1"])


def test_line_forward(golden):
    g = golden
    x = g["line/x"]
    for bits in (8, 4):
        for fzp in (True, False):
            y = orc.fq_line_fwd(x, np.array([[-0.1, 0.9]], np.float32), bits, -1, fzp)
            assert bits_equal(y, g[f"line/y_tuple_b{bits}_fzp{int(fzp)}"]), (bits, fzp)
            y = orc.fq_line_fwd(x, g["line/lines_ch1"], bits, 1, fzp)
            assert bits_equal(y, g[f"line/y_ch1_b{bits}_fzp{int(fzp)}"]), (bits, fzp)


def test_ste_backward(golden):
    g = golden
    gr = g["bwd/g"]
    for bits, d, flip in ((8, 5, False), (8, 5, True), (4, 3, False), (8, -1, False)):
        gc, _ = orc.ste_bwd(gr, np.float32(d), bits, -1, True, flip)
        assert bits_equal(gc, g[f"bwd/pow2_b{bits}_d{d}_f{int(flip)}_gx"])
        # the reference clamps grad_output in place (quantize.py:72)
        assert bits_equal(gc, g[f"bwd/pow2_b{bits}_d{d}_f{int(flip)}_go"])
    gc, _ = orc.ste_bwd(gr, g["pow2/dec_ch1"], 8, 1, True, False)
    assert bits_equal(gc, g["bwd/pow2_ch1_gx"])
    assert bits_equal(gr, g["bwd/pow2_passthrough_gx"])
    for name, s in (("0p1", 0.1), ("0p037", 0.037)):
        gc, _ = orc.ste_bwd(gr, np.float32(s), 8, -1, False, False)
        assert bits_equal(gc, g[f"bwd/scaler_s{name}_gx"])
    gc, _ = orc.ste_bwd(gr, g["scaler/s_ch1"], 6, 1, False, True)
    assert bits_equal(gc, g["bwd/scaler_ch1_b6_flip_gx"])


def test_decimal_quantizer_optimize(golden):
    g = golden
    xs = g["dq/xs"]
    for bits in (8, 4):
        w = np.zeros(1, np.float32)
        for t, x in enumerate(xs):
            w = orc.scale_ema(w, orc.absmax(x, -1), bits, t)
            assert bits_equal(w, g[f"dq/w_tensor_b{bits}"][t].reshape(-1)), (bits, t)
    for ci, key in ((0, "dq/w_ch0"), (1, "dq/w_ch1")):
        w = np.zeros(xs.shape[1 + ci], np.float32)
        for t, x in enumerate(xs):
            w = orc.scale_ema(w, orc.absmax(x, ci), 8, t)
            assert bits_equal(w, g[key][t].reshape(-1)), (ci, t)


def test_scale_to_decimal(golden):
    g = golden
    with np.errstate(all="ignore"):
        d = orc.scale_to_decimal(g["dq/scales"])
    assert bits_equal(d, g["dq/decimals"])
    w = g["dq/w_tensor_b8"][0].reshape(-1)
    y = orc.fq_pow2_fwd(g["dq/xs"][0], orc.scale_to_decimal(w))
    assert bits_equal(y, g["dq/fwd_tensor"])


def test_adaptive_optimize(golden):
    g = golden
    xs = g["dq/xs"]
    for ci, batched, nch in ((1, True, 4), (0, False, 2), (-1, True, 1)):
        lines = np.zeros((nch, 2), np.float32)
        for t, x in enumerate(xs):
            mn, mx = orc.minmax(x, ci)
            lines = orc.lines_ema(lines, mn, mx, t + 1)
            assert bits_equal(lines, g[f"aq/lines_ci{ci}_b{int(batched)}"][t]), (ci, t)


def test_squeeze_mean_abs(golden):
    g = golden
    x = g["sq/x"]
    for tgt in ((1, 8, 1, 1), (1, 8, 5, 7), (6, 1, 1, 1), (1, 8, 5, 1), (6, 8, 5, 7), (1, 1, 5, 7)):
        ref = g["sq/" + "x".join(map(str, tgt))]
        got = orc.squeeze_mean_abs(x, tgt)
        assert got.shape == ref.shape
        # torch's fp32 cascade summation is not restated: a few ulp
        assert ulp_diff(got, ref).max() <= 4, tgt


def test_mask_given_importance(golden):
    g = golden
    for s in (0.0, 0.47, 0.5, 0.75, 0.999):
        m, _ = orc.mask_given_importance(g["mask/imp"], s)
        assert np.array_equal(m, g[f"mask/m_{s}"]), s
    assert 1 - orc.mask_given_importance(g["mask/imp"], 0.47)[0].mean() == 0.47  # tests/test_util.py:80-84
    for s in (0.0, 0.25, 0.5, 0.75):
        m, _ = orc.mask_given_importance(g["mask/tie_imp"], s)
        assert np.array_equal(m, g[f"mask/tie_m_{s}"]), s
    m, _ = orc.mask_given_importance(g["mask/neg_imp"], 0.3)
    assert np.array_equal(m, g["mask/neg_m_0.3"])


def _replay_callback(g, tag, mask_shape, sparsity=0.5, running_average=True, interval=1, stop=float("inf")):
    xs, outs, masks, mags = g[f"cb/{tag}_x"], g[f"cb/{tag}_out"], g[f"cb/{tag}_mask"], g[f"cb/{tag}_mag"]
    mask = np.ones(mask_shape, bool)
    mag = np.zeros(mask_shape, np.float32)
    structured = int(np.prod(mask_shape)) != xs[0].size
    for t, x in enumerate(xs):
        if t < stop and running_average:
            mag = orc.magnitude_ema(mag, orc.squeeze_mean_abs(x, mask_shape), t)
            assert ulp_diff(mag, mags[t]).max() <= 8, (tag, t)
        if orc.refresh_gate(t, sparsity, interval, stop, running_average):
            imp = mag if running_average else orc.squeeze_mean_abs(x, mask_shape)
            mask, _ = orc.mask_given_importance(imp, sparsity)
        assert np.array_equal(mask, masks[t]), (tag, t)
        out = orc.mask_apply(x, mask.reshape(-1), 1 if structured else -1)
        assert bits_equal(out, outs[t]), (tag, t)


def test_magnitude_callback_sequences(golden):
    _replay_callback(golden, "struct", (1, 8, 1, 1))
    _replay_callback(golden, "unstruct", (2, 3, 6, 6))
    _replay_callback(golden, "struct_norunavg", (1, 8, 1, 1), running_average=False)
    _replay_callback(golden, "struct_refresh2", (1, 8, 1, 1), interval=2, stop=4)


def test_prune_ramp(golden):
    ref = golden["ramp/cur_sparsity"]
    schedules, rampup_interval = orc.prune_schedule(200, 10, 4)
    cur = 0.0
    for step in range(241):
        if step in schedules:
            cur = orc.ramp_sparsity(step, 0.5, 200, 10, 4, rampup_interval)
        assert cur == ref[step], step
    # docs/tutorial.ipynb:224-228 : 0.29 / 0.44 / 0.49 / 0.50
    assert [round(ref[s], 2) for s in (200, 210, 220, 230)] == [0.29, 0.44, 0.49, 0.5]


def test_oracle_layer_emulators_match_reference_layers(golden):
    """tests/oracle_layers.py (host control flow of QuantizeLayer / PruneLayer over the oracle) against the
    reference's recorded layer flows — pins the emulators used for the teacher-forced end-to-end GPU test."""
    from tests.oracle_layers import OraclePrune, OracleQuantize
    g = golden
    x = g["layer/x"]
    for kind, key in (("decimal", "layer/q_dec"), ("scaler", "layer/q_scl")):
        em = OracleQuantize(8, 3, kind)
        outs = np.stack([em.forward(x) for _ in range(6)])
        assert bits_equal(outs, g[key + "_out"]) and bits_equal(em.weight, g[key + "_weight"].reshape(-1))
    em = OraclePrune(0.5, 2, 2, 3)
    for t in range(12):
        out = em.forward(g["layer/px"])
        assert bits_equal(out, g["layer/p_out"][t]), t
        assert np.array_equal(em.mask, g["layer/p_mask"][t]), t


DIMS_CASES = {"nchw_13": {1, 3}, "nchw_03": {0, 3}, "w_02": {0, 2}, "w_013": {0, 1, 3}, "nlc_02": {0, 2}}


@pytest.mark.parametrize("name", sorted(DIMS_CASES))
def test_non_adjacent_kept_axes(name):
    """Prune masks whose kept axes are not adjacent (dimensions={1,3} on NCHW ...), recorded from the reference by
    oracle/gen_golden_dims.py: the oracle's squeeze / EMA / mask flow reproduces magnitudes (<= 4 ulp, the fp32
    mean chain is not restated), masks and outputs."""
    from pathlib import Path
    from tests.oracle_layers import OraclePrune
    g = np.load(Path(__file__).resolve().parent / "golden" / "dims_v1.npz")
    em = OraclePrune(0.5, 2, 2, 2, dimensions=DIMS_CASES[name])
    for t in range(g[f"{name}/x"].shape[0]):
        out = em.forward(g[f"{name}/x"][t])
        assert np.array_equal(em.mask.reshape(g[f"{name}/mask"][t].shape), g[f"{name}/mask"][t]), t
        assert bits_equal(out, g[f"{name}/out"][t]), t
        if em.mag is not None:
            assert ulp_diff(em.mag.reshape(-1), g[f"{name}/mag"][t].reshape(-1)).max() <= 4, t


# ----------------------------------------------------------------------------- corners (oracle/gen_golden_extremes.py)
@pytest.fixture(scope="module")
def extremes():
    from pathlib import Path
    return np.load(Path(__file__).resolve().parent / "golden" / "extremes_v1.npz")


def test_extreme_parameters_forward(extremes):
    """Overflowing decimals, zero / negative / infinite / NaN scales, empty and inverted line ranges over +-0, +-inf,
    NaN, subnormal and huge inputs: the oracle equals the reference everywhere, including where `.int()` is out of
    range (x86 semantics: INT_MIN)."""
    g = extremes
    x = g["x"]
    assert bits_equal(orc.fq_pow2_fwd(x, g["decs"], 1), g["pow2/ch"])
    assert bits_equal(orc.fq_scaler_fwd(x, g["scales"], 1), g["scaler/ch"])
    for i, d in enumerate(g["decs"]):
        assert bits_equal(orc.fq_pow2_fwd(x, float(d)), g[f"pow2/t{i}"]), d
    for i, s in enumerate(g["scales"]):
        assert bits_equal(orc.fq_scaler_fwd(x, np.float32(s)), g[f"scaler/t{i}"]), s
    for fzp in (True, False):
        assert bits_equal(orc.fq_line_fwd(x, g["lines"], 8, 1, fzp), g[f"line/ch_fzp{int(fzp)}"]), fzp
        for i, ln in enumerate(g["lines"]):
            assert bits_equal(orc.fq_line_fwd(x, ln.reshape(1, 2), 8, -1, fzp), g[f"line/t{i}_fzp{int(fzp)}"]), (i, fzp)


@pytest.mark.parametrize("bits", [1, 2, 3, 12, 16, 24, 32])
def test_bit_width_extremes(extremes, bits):
    g = extremes
    for fzp in (True, False):
        assert bits_equal(orc.fq_line_fwd(g["bits/x"], g["bits/lines"], bits, 1, fzp),
                          g[f"bits/line_b{bits}_fzp{int(fzp)}"]), fzp
    for name, par, is_dec in (("dec", g["bits/dec"], True), ("scale", g["bits/scale"], False)):
        for flip in (False, True):
            got, _ = orc.ste_bwd(g["bits/g"], par, bits, 1, is_dec, flip)
            assert bits_equal(got, g[f"bits/bwd_{name}_b{bits}_f{int(flip)}"]), (name, flip)


def test_mask_corner_cases(extremes):
    """NaN / inf / constant / tiny importances at sparsity 0 ... 1: same mask, or IndexError where the reference
    raises it."""
    g = extremes
    for name in g["mask/names"]:
        imp = g[f"mask/{name}/imp"]
        for i, sp in enumerate(g["mask/sparsities"]):
            exp = g[f"mask/{name}/s{i}"]
            if exp.dtype == np.int8:
                with pytest.raises(IndexError):
                    orc.mask_given_importance(imp, float(sp))
            else:
                assert np.array_equal(orc.mask_given_importance(imp, float(sp))[0], exp), (name, sp)


def test_scale_to_decimal_degenerate_scales(extremes):
    """zero / infinite / NaN / negative / subnormal scales: 1/s overflows to the nan_to_num replacement values."""
    assert bits_equal(orc.scale_to_decimal(extremes["s2d/scales"]), extremes["s2d/decimals"])
