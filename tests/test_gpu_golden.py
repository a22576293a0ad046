"""GPU parity, part 1: the product API (qsparse_b200, CUDA kernels through the C-ABI)
against golden vectors recorded from the reference itself (tests/golden/, generated
by oracle/gen_golden.py).  Bit-exact unless a tolerance is stated."""
import contextlib
import io
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.conftest import bits_equal, ulp_diff

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def q():
    import importlib
    import qsparse_b200
    qsparse_b200.set_qsparse_options(log_on_created=False)
    # `qsparse_b200.quantize` the attribute is the quantize() function (as in the
    # reference's __init__), so fetch the submodule explicitly
    return importlib.import_module("qsparse_b200.quantize")


def test_pow2_forward(golden, q):
    g = golden
    x = cu(g["pow2/x"])
    for d in (5, 0, -2, 12):
        assert bits_equal(npy(q.quantize_with_decimal(x, 8, d)), g[f"pow2/y_d{d}"]), d
    assert bits_equal(npy(q.quantize_with_decimal(x, 8, cu(g["pow2/dec_ch1"]), 1)), g["pow2/y_ch1"])
    assert bits_equal(npy(q.quantize_with_decimal(x, 8, cu(g["pow2/dec_ch0"]), 0)), g["pow2/y_ch0"])
    assert bits_equal(npy(q.quantize_with_decimal(x, 4, 5, -1, True, False, True)), g["pow2/y_d5_uint_flip_b4"])
    assert bits_equal(npy(q.quantize_with_decimal(cu(g["pow2/x2"]), 8, cu(g["pow2/dec2"]), 1)), g["pow2/y2_ch1"])
    with pytest.raises(AssertionError):
        q.quantize_with_decimal(x, 8, cu(g["pow2/dec_ch0"]), 1)  # channel count mismatch (quantize.py:49-51)


def test_scaler_forward(golden, q):
    g = golden
    x = cu(g["scaler/x"])
    for name, s in (("0p1", 0.1), ("0p037", 0.037), ("0p5", 0.5), ("3", 3.0)):
        assert bits_equal(npy(q.quantize_with_scaler(x, 8, s)), g[f"scaler/y_s{name}"]), name
    assert bits_equal(npy(q.quantize_with_scaler(x, 8, cu(g["scaler/s_ch1"]), 1)), g["scaler/y_ch1"])


def test_line_forward(golden, q):
    g = golden
    x = cu(g["line/x"])
    for bits in (8, 4):
        for fzp in (True, False):
            y = q.quantize_with_line(x, bits, (-0.1, 0.9), -1, False, fzp)
            assert bits_equal(npy(y), g[f"line/y_tuple_b{bits}_fzp{int(fzp)}"]), (bits, fzp)
            y = q.quantize_with_line(x, bits, cu(g["line/lines_ch1"]), 1, False, fzp)
            assert bits_equal(npy(y), g[f"line/y_ch1_b{bits}_fzp{int(fzp)}"]), (bits, fzp)


def _run_bwd(fn, x_np, g_np, *args):
    x = cu(x_np).requires_grad_(True)
    y = fn(x, *args)
    go = cu(g_np).clone()
    y.backward(go)
    return npy(x.grad), npy(go)


def test_ste_backward(golden, q):
    g = golden
    x, gr = g["pow2/x"], g["bwd/g"]
    for bits, d, flip in ((8, 5, False), (8, 5, True), (4, 3, False), (8, -1, False)):
        gx, go = _run_bwd(q.quantize_with_decimal, x, gr, bits, d, -1, False, False, flip)
        assert bits_equal(gx, g[f"bwd/pow2_b{bits}_d{d}_f{int(flip)}_gx"])
        assert bits_equal(go, g[f"bwd/pow2_b{bits}_d{d}_f{int(flip)}_go"])  # in-place clamp of grad_output
    gx, go = _run_bwd(q.quantize_with_decimal, x, gr, 8, cu(g["pow2/dec_ch1"]), 1, False, False, False)
    assert bits_equal(gx, g["bwd/pow2_ch1_gx"]) and bits_equal(go, g["bwd/pow2_ch1_go"])
    gx, _ = _run_bwd(q.quantize_with_decimal, x, gr, 8, 5, -1, False, True, False)
    assert bits_equal(gx, g["bwd/pow2_passthrough_gx"])
    for name, s in (("0p1", 0.1), ("0p037", 0.037)):
        gx, _ = _run_bwd(q.quantize_with_scaler, x, gr, 8, s, -1, False, False, False)
        assert bits_equal(gx, g[f"bwd/scaler_s{name}_gx"])
    gx, _ = _run_bwd(q.quantize_with_scaler, x, gr, 6, cu(g["scaler/s_ch1"]), 1, False, False, True)
    assert bits_equal(gx, g["bwd/scaler_ch1_b6_flip_gx"])
    # LineQuantization: identity backward (quantize.py:183-185)
    gx, go = _run_bwd(q.quantize_with_line, x, gr, 8, (-0.1, 0.9))
    assert bits_equal(gx, gr)


def test_decimal_quantizer_optimize(golden, q):
    g = golden
    xs = g["dq/xs"]
    for bits in (8, 4):
        cb = q.DecimalQuantizer()
        w = torch.zeros(1, 1, device="cuda")
        for t, x in enumerate(xs):
            w.data[:] = cb.optimize(cu(x), bits, w, batched=True, channel_index=-1)
            assert bits_equal(npy(w), g[f"dq/w_tensor_b{bits}"][t]), (bits, t)
    for ci, key in ((0, "dq/w_ch0"), (1, "dq/w_ch1")):
        cb = q.ScalerQuantizer()
        w = torch.zeros(xs.shape[1 + ci], 1, device="cuda")
        for t, x in enumerate(xs):
            w.data[:] = cb.optimize(cu(x), 8, w, batched=False, channel_index=ci)
            assert bits_equal(npy(w), g[key][t]), (ci, t)
    with pytest.raises(RuntimeError):  # the reference raises for batch > 1
        q.DecimalQuantizer().optimize(cu(xs[0]), 8, torch.zeros(4, 1, device="cuda"), batched=True, channel_index=1)


def test_scale_to_decimal(golden, q):
    from qsparse_b200 import ops
    g = golden
    d = ops.scale_to_decimal(cu(g["dq/scales"]))
    assert bits_equal(npy(d), g["dq/decimals"])
    cb = q.DecimalQuantizer()
    w = torch.zeros(1, 1, device="cuda")
    x0 = cu(g["dq/xs"][0])
    w.data[:] = cb.optimize(x0, 8, w, batched=True, channel_index=-1)
    assert bits_equal(npy(cb(x0, 8, w, channel_index=-1)), g["dq/fwd_tensor"])


def test_adaptive_optimize(golden, q):
    g = golden
    xs = g["dq/xs"]
    for ci, batched, nch in ((1, True, 4), (0, False, 2), (-1, True, 1)):
        cb = q.AdaptiveQuantizer()
        w = torch.zeros(nch, 2, device="cuda")
        for t, x in enumerate(xs):
            w.data[:] = cb.optimize(cu(x), 8, w, channel_index=ci, batched=batched)
            assert bits_equal(npy(w), g[f"aq/lines_ci{ci}_b{int(batched)}"][t]), (ci, t)


def test_squeeze_tensor_to_shape(golden):
    from qsparse_b200.util import squeeze_tensor_to_shape, mean_abs_to_shape
    g = golden
    x = cu(g["sq/x"])
    for tgt in ((1, 8, 1, 1), (1, 8, 5, 7), (6, 1, 1, 1), (1, 8, 5, 1), (6, 8, 5, 7), (1, 1, 5, 7)):
        ref = g["sq/" + "x".join(map(str, tgt))]
        got = npy(squeeze_tensor_to_shape(x.abs(), tgt))
        assert got.shape == ref.shape
        assert ulp_diff(got, ref).max() <= 4, tgt  # fp32 cascade sums are not restated
        got = npy(mean_abs_to_shape(x, tgt))
        assert ulp_diff(got, ref).max() <= 4, tgt
    # kept axes 0 and 2 are not adjacent: the transposing route (ref: any dimension set, util.py:93-101)
    want = npy(x.abs().mean(dim=(1, 3), keepdim=True))
    assert ulp_diff(npy(squeeze_tensor_to_shape(x.abs(), (6, 1, 5, 1))), want).max() <= 4
    assert ulp_diff(npy(mean_abs_to_shape(x, (6, 1, 5, 1))), want).max() <= 4
    with pytest.raises(ValueError):
        squeeze_tensor_to_shape(x.abs(), (1, 3, 1, 1))


def test_mask_given_importance(golden):
    from qsparse_b200 import calculate_mask_given_importance
    g = golden
    imp = cu(g["mask/imp"])
    for s in (0.0, 0.47, 0.5, 0.75, 0.999):
        assert np.array_equal(npy(calculate_mask_given_importance(imp, s)), g[f"mask/m_{s}"]), s
    m = calculate_mask_given_importance(imp, 0.47)
    assert (1 - m.sum().item() / m.numel()) == 0.47  # reference tests/test_util.py:80-84
    for s in (0.0, 0.25, 0.5, 0.75):
        assert np.array_equal(npy(calculate_mask_given_importance(cu(g["mask/tie_imp"]), s)), g[f"mask/tie_m_{s}"])
    assert np.array_equal(npy(calculate_mask_given_importance(cu(g["mask/neg_imp"]), 0.3)), g["mask/neg_m_0.3"])


def _replay_callback(g, tag, mask_shape, **kw):
    from qsparse_b200 import MagnitudePruningCallback
    xs, outs, masks, mags = g[f"cb/{tag}_x"], g[f"cb/{tag}_out"], g[f"cb/{tag}_mask"], g[f"cb/{tag}_mag"]
    cb = MagnitudePruningCallback(**kw).cuda()
    cb.train()
    mask = torch.nn.Parameter(torch.ones(*mask_shape, dtype=torch.bool, device="cuda"), requires_grad=False)
    for t, x in enumerate(xs):
        out = cb(cu(x), 0.5, mask)
        assert np.array_equal(npy(mask), masks[t]), (tag, t)
        assert bits_equal(npy(out), outs[t]), (tag, t)
        if mags.size:
            assert ulp_diff(npy(cb.magnitude), mags[t]).max() <= 8, (tag, t)
    assert cb.t.item() == len(xs)


def test_magnitude_callback_sequences(golden):
    _replay_callback(golden, "struct", (1, 8, 1, 1))
    _replay_callback(golden, "unstruct", (2, 3, 6, 6))
    _replay_callback(golden, "struct_norunavg", (1, 8, 1, 1), running_average=False)
    _replay_callback(golden, "struct_refresh2", (1, 8, 1, 1), mask_refresh_interval=2, stop_mask_refresh=4)


def test_prune_layer_ramp(golden):
    from qsparse_b200.sparse import PruneLayer, MagnitudePruningCallback
    ref = golden["ramp/cur_sparsity"]
    layer = PruneLayer(sparsity=0.5, start=200, interval=10, repetition=4, dimensions={1},
                       callback=MagnitudePruningCallback())
    layer.train()
    rng = np.random.default_rng(1)
    with contextlib.redirect_stdout(io.StringIO()):
        for step in range(241):
            layer(cu(rng.random((2, 8, 3, 3)).astype(np.float32)))
            if step in (0, 199, 200, 205, 210, 220, 230, 240):
                assert layer._cur_sparsity.item() == ref[step], step
    assert layer._n_updates.item() == 241 and layer._n_updates.dtype == torch.int32
    assert int((~layer.mask).sum().item()) == 4  # 50 % of 8 channels


def test_layer_flows(golden):
    import qsparse_b200 as qs
    from qsparse_b200.quantize import DecimalQuantizer, AdaptiveQuantizer, ScalerQuantizer
    from qsparse_b200.sparse import MagnitudePruningCallback
    g = golden
    data = cu(g["layer/x"])
    with contextlib.redirect_stdout(io.StringIO()):
        ql = qs.quantize(bits=8, timeout=3, channelwise=-1, callback=DecimalQuantizer())
        outs = np.stack([npy(ql(data)) for _ in range(6)])
        assert bits_equal(outs, g["layer/q_dec_out"]) and bits_equal(npy(ql.weight), g["layer/q_dec_weight"])
        ql = qs.quantize(bits=8, timeout=3, channelwise=1, callback=AdaptiveQuantizer())
        outs = np.stack([npy(ql(data)) for _ in range(6)])
        assert bits_equal(outs, g["layer/q_adp_out"]) and bits_equal(npy(ql.weight), g["layer/q_adp_weight"])
        ql.eval()
        assert bits_equal(npy(ql(data)), g["layer/q_adp_eval_out"])
        ql = qs.quantize(bits=8, timeout=3, channelwise=-1)
        outs = np.stack([npy(ql(data)) for _ in range(6)])
        assert bits_equal(outs, g["layer/q_scl_out"]) and bits_equal(npy(ql.weight), g["layer/q_scl_weight"])
        assert sorted(ql.state_dict().keys()) == ["_n_updates", "weight"]

        pl = qs.prune(sparsity=0.5, start=2, interval=2, repetition=3, dimensions={1})
        big = cu(g["layer/px"])
        for t in range(12):
            out = pl(big)
            assert np.array_equal(npy(pl.mask), g["layer/p_mask"][t]), t
            assert bits_equal(npy(out), g["layer/p_out"][t]), t
        assert sorted(pl.state_dict().keys()) == ["_cur_sparsity", "_n_updates", "callback.magnitude", "callback.t",
                                                  "mask"]

        conv = torch.nn.Conv2d(4, 6, 3)
        with torch.no_grad():
            conv.weight.copy_(torch.from_numpy(g["layer/conv_w"]))
            conv.bias.copy_(torch.from_numpy(g["layer/conv_b"]))
        conv = conv.cuda()
        qp = qs.quantize(qs.prune(conv, sparsity=0.5, start=1, interval=1, repetition=2,
                                  callback=MagnitudePruningCallback(running_average=False)),
                         bits=8, timeout=2, channelwise=0, callback=ScalerQuantizer())
        qp.train()
        for t in range(6):
            assert bits_equal(npy(qp.weight), g["layer/qp_weight_seq"][t]), t
        assert np.array_equal(npy(qp.prune.mask), g["layer/qp_mask"])
        assert bits_equal(npy(qp.quantize.weight), g["layer/qp_scale"])
        assert not bits_equal(npy(qp._parameters["weight"]), npy(qp.weight))  # raw parameter untouched


DIMS_CASES = {"nchw_13": ({1, 3}, True), "nchw_03": ({0, 3}, True), "w_02": ({0, 2}, True),
              "w_02_instant": ({0, 2}, False), "w_013": ({0, 1, 3}, True), "nlc_02": ({0, 2}, True)}


@pytest.mark.parametrize("name", sorted(DIMS_CASES))
def test_non_adjacent_kept_axes(name):
    """PruneLayer with a dimension set whose axes are not adjacent (mask [1,C,1,W], [Cout,1,kh,1] ...) against
    what the reference returned (oracle/gen_golden_dims.py): outputs, masks, input gradients bit for bit, running
    magnitudes to 4 ulp."""
    import qsparse_b200 as qs
    from qsparse_b200.sparse import MagnitudePruningCallback
    g = np.load(Path(__file__).resolve().parent / "golden" / "dims_v1.npz")
    dims, ra = DIMS_CASES[name]
    with contextlib.redirect_stdout(io.StringIO()):
        pl = qs.prune(sparsity=0.5, start=2, interval=2, repetition=2, dimensions=dims,
                      callback=MagnitudePruningCallback(running_average=ra))
        pl.train()
        for t in range(g[f"{name}/x"].shape[0]):
            x = cu(g[f"{name}/x"][t]).requires_grad_(True)
            y = pl(x)
            y.backward(cu(g[f"{name}/g"][t]))
            assert np.array_equal(npy(pl.mask), g[f"{name}/mask"][t]), t
            assert bits_equal(npy(y), g[f"{name}/out"][t]), t
            assert bits_equal(npy(x.grad), g[f"{name}/gx"][t]), t
            if ra and hasattr(pl.callback, "magnitude"):
                assert ulp_diff(npy(pl.callback.magnitude).reshape(-1), g[f"{name}/mag"][t].reshape(-1)).max() <= 4, t


# ----------------------------------------------------------------------------- corners (oracle/gen_golden_extremes.py)
@pytest.fixture(scope="module")
def extremes():
    from pathlib import Path
    return np.load(Path(__file__).resolve().parent / "golden" / "extremes_v1.npz")


def test_extreme_parameters_against_reference_goldens(extremes, q):
    """The kernels against what the reference itself returned for overflowing decimals, degenerate scales and line
    ranges over non-finite inputs.  The line quantizer has no int32 step: compared everywhere.  Decimal / scaler go
    through `.int()`, which is INT_MIN on the reference's x86 run and saturating on CUDA (for the reference as
    well): compared inside the documented domain |q| < 2^31."""
    g = extremes
    x = g["x"]
    xc = cu(x)
    C = x.shape[1]
    for fzp in (True, False):
        assert bits_equal(npy(q.quantize_with_line(xc, 8, cu(g["lines"]), 1, False, fzp)), g[f"line/ch_fzp{int(fzp)}"])
        for i, ln in enumerate(g["lines"]):
            got = q.quantize_with_line(xc, 8, (float(ln[0]), float(ln[1])), -1, False, fzp)
            assert bits_equal(npy(got), g[f"line/t{i}_fzp{int(fzp)}"]), (i, fzp)

    def inside(qv):
        return np.isfinite(qv) & (np.abs(qv) < 2.0 ** 31)

    with np.errstate(all="ignore"):
        ok = inside(x * (np.float32(2.0) ** g["decs"]).reshape(1, C, 1))
        got = npy(q.quantize_with_decimal(xc, 8, cu(g["decs"]), 1))
        assert ok.mean() > 0.3 and bits_equal(np.where(ok, got, 0), np.where(ok, g["pow2/ch"], 0))
        ok = inside(x / g["scales"].reshape(1, C, 1))
        got = npy(q.quantize_with_scaler(xc, 8, cu(g["scales"]), 1))
        assert ok.mean() > 0.3 and bits_equal(np.where(ok, got, 0), np.where(ok, g["scaler/ch"], 0))
        for i, d in enumerate(g["decs"]):
            ok = inside(x * np.float32(2.0) ** d)
            got = npy(q.quantize_with_decimal(xc, 8, float(d)))
            assert bits_equal(np.where(ok, got, 0), np.where(ok, g[f"pow2/t{i}"], 0)), d
        for i, s in enumerate(g["scales"]):
            if np.isnan(s):
                continue
            ok = inside(x / s)
            got = npy(q.quantize_with_scaler(xc, 8, float(s)))
            assert bits_equal(np.where(ok, got, 0), np.where(ok, g[f"scaler/t{i}"], 0)), s


@pytest.mark.parametrize("bits", [1, 2, 3, 12, 16, 24, 32])
def test_bit_width_extremes_against_reference_goldens(extremes, q, bits):
    """1 ... 32 bits through the public functions: line forward, and the in-place clamp of grad_output by the
    straight-through backward (at 1 bit a bound is +0.0: the reference keeps a -0.0 gradient, FMNMX would not)."""
    g = extremes
    x, gr = g["bits/x"], g["bits/g"]
    for fzp in (True, False):
        got = q.quantize_with_line(cu(x), bits, cu(g["bits/lines"]), 1, False, fzp)
        assert bits_equal(npy(got), g[f"bits/line_b{bits}_fzp{int(fzp)}"]), fzp
    for name, fn, par in (("dec", q.quantize_with_decimal, g["bits/dec"]),
                          ("scale", q.quantize_with_scaler, g["bits/scale"])):
        for flip in (False, True):
            _, go = _run_bwd(fn, x, gr, bits, cu(par), 1, False, False, flip)
            assert bits_equal(go, g[f"bits/bwd_{name}_b{bits}_f{int(flip)}"]), (name, flip)


def test_mask_corner_cases_against_reference_goldens(extremes):
    from qsparse_b200.util import calculate_mask_given_importance
    g = extremes
    for name in g["mask/names"]:
        imp = cu(g[f"mask/{name}/imp"])
        for i, sp in enumerate(g["mask/sparsities"]):
            exp = g[f"mask/{name}/s{i}"]
            if exp.dtype == np.int8:
                with pytest.raises(IndexError):
                    calculate_mask_given_importance(imp, float(sp))
            else:
                assert np.array_equal(npy(calculate_mask_given_importance(imp, float(sp))), exp), (name, sp)


def test_scale_to_decimal_degenerate_scales_against_reference_goldens(extremes):
    from qsparse_b200 import ops
    got = ops.scale_to_decimal(cu(extremes["s2d/scales"]))
    assert bits_equal(npy(got).reshape(-1), extremes["s2d/decimals"])
