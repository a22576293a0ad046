"""Host-side logic of the quantize/prune layers (no kernels): schedules, layout
factorisation, host mirrors, conversion structure, API surface."""
import contextlib
import io

import numpy as np
import pytest
import torch
import torch.nn as nn

import qsparse_b200 as qs
from qsparse_b200 import ops
from qsparse_b200._native import channel_layout
from qsparse_b200.util import HostMirror, kth_rank
from oracle import oracle as orc


def setup_module(_):
    qs.set_qsparse_options(log_on_created=False)


def test_api_surface_matches_reference_exports():
    # qsparse/__init__.py:2-8
    for name in ("convert", "quantize", "DecimalQuantizer", "ScalerQuantizer", "AdaptiveQuantizer",
                 "MagnitudePruningCallback", "UniformPruningCallback", "prune", "devise_layerwise_pruning_schedule",
                 "auto_name_prune_quantize_layers", "calculate_mask_given_importance", "get_qsparse_option",
                 "set_qsparse_options"):
        assert hasattr(qs, name), name
    import importlib
    q = importlib.import_module("qsparse_b200.quantize")
    for name in ("quantize_with_decimal", "quantize_with_scaler", "quantize_with_line", "QuantizeLayer",
                 "BaseQuantizer", "DecimalQuantization", "ScalerQuantization", "LineQuantization"):
        assert hasattr(q, name), name
    u = importlib.import_module("qsparse_b200.util")
    for name in ("squeeze_tensor_to_shape", "preload_qsparse_state_dict", "nn_module"):
        assert hasattr(u, name), name


def test_every_reference_export_is_importable_and_alias_works():
    """qsparse/__init__.py:2-8 of the reference, name by name (fuse_bn included), and the `qsparse` alias"""
    import importlib
    import sys
    for name in ("convert", "fuse_bn", "quantize", "DecimalQuantizer", "ScalerQuantizer", "AdaptiveQuantizer",
                 "MagnitudePruningCallback", "UniformPruningCallback", "prune", "devise_layerwise_pruning_schedule",
                 "auto_name_prune_quantize_layers", "calculate_mask_given_importance", "get_qsparse_option",
                 "set_qsparse_options"):
        assert getattr(qs, name) is not None, name
    saved = {k: v for k, v in sys.modules.items() if k == "qsparse" or k.startswith("qsparse.")}
    for k in saved:
        del sys.modules[k]
    try:
        from qsparse_b200 import compat
        compat.install_as_qsparse()
        qsparse = importlib.import_module("qsparse")
        assert qsparse is qs and qsparse.fuse_bn is qs.fuse_bn
        from qsparse.quantize import QuantizeLayer, quantize_with_decimal   # noqa: F401
        from qsparse.sparse import PruneLayer                               # noqa: F401
        from qsparse.util import squeeze_tensor_to_shape                    # noqa: F401
        from qsparse.fuse import fuse_bn                                    # noqa: F401
    finally:
        for k in [k for k in sys.modules if k == "qsparse" or k.startswith("qsparse.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_channel_layout():
    assert channel_layout((256, 64, 56, 56), 1) == (256, 64, 3136)
    assert channel_layout((4096, 4096), 0) == (1, 4096, 4096)
    assert channel_layout((8, 16), 1) == (8, 16, 1)
    assert channel_layout((8, 16, 3, 3), -1) == (1, 1, 8 * 16 * 9)


def test_mask_layout():
    assert ops.mask_layout((4, 8, 5, 5), (1, 8, 1, 1)) == ("channel", (4, 8, 25))
    assert ops.mask_layout((4, 8, 5, 5), (4, 8, 5, 5)) == ("element", (1, 1, 800))
    assert ops.mask_layout((4, 8, 5, 5), (1, 8, 5, 5)) == ("channel", (4, 200, 1))
    assert ops.mask_layout((6, 4, 3, 3), (6, 1, 1, 1)) == ("channel", (1, 6, 36))
    assert ops.mask_layout((4, 8, 10, 10), (1, 1, 1, 1)) == ("channel", (1, 1, 3200))
    assert ops.mask_layout((1, 8, 5, 5), (1, 8, 5, 5)) == ("element", (1, 1, 200))
    with pytest.raises(RuntimeError):      # unstructured mask meets a new spatial size (tests/test_sparse.py:75-77)
        ops.mask_layout((1, 10, 64, 64), (1, 10, 32, 32))
    with pytest.raises(NotImplementedError):
        ops.mask_layout((4, 8, 5, 5), (4, 1, 5, 1))
    # ... which is what mask_perm() resolves: move the sandwiched broadcast axes in front of the kept run
    assert ops.mask_perm((4, 8, 5, 5), (1, 8, 1, 1)) is None and ops.mask_perm((4, 8, 5, 5), (4, 8, 5, 5)) is None
    perm = ops.mask_perm((4, 8, 5, 6), (4, 1, 5, 1))
    assert perm == [1, 0, 2, 3]
    assert ops.mask_layout([(4, 8, 5, 6)[p] for p in perm], [(4, 1, 5, 1)[p] for p in perm]) == ("channel", (8, 20, 6))
    perm = ops.mask_perm((4, 8, 5, 6), (1, 8, 1, 6))
    assert perm == [0, 2, 1, 3] and ops.invert_perm(perm) == [0, 2, 1, 3]
    assert ops.mask_layout([(4, 8, 5, 6)[p] for p in perm], [(1, 8, 1, 6)[p] for p in perm]) == ("channel", (20, 48, 1))
    perm = ops.mask_perm((2, 3, 4, 5, 6), (2, 1, 4, 1, 6))
    assert perm == [1, 3, 0, 2, 4] and ops.invert_perm(perm) == [2, 0, 3, 1, 4]


def test_kth_rank_matches_reference_formula():
    for s in (0.0, 0.1, 0.47, 0.5, 0.75, 0.999, 1.0):
        for n in (2, 8, 64, 1000, 16800):
            assert kth_rank(s, n) == orc.kth_index(s, n) == max(int(s * n - 1), 0) + 1


def test_prune_layer_schedule_and_ramp(golden):
    """PruneLayer's host arithmetic: schedules, cubic ramp, fp32 round trip (sparse.py:186-195,252-257)."""
    from qsparse_b200.sparse import PruneLayer
    layer = PruneLayer(sparsity=0.5, start=200, interval=10, repetition=4)
    assert layer.schedules == [200, 210, 220, 230] and layer.rampup_interval == 10
    layer = PruneLayer(sparsity=0.5, start=200, interval=10, repetition=4, rampup=True)
    assert layer.schedules == [210, 220, 230, 240] and layer.rampup_interval == 0
    ref = golden["ramp/cur_sparsity"]
    for n in (200, 210, 220, 230):
        ratio = (1.0 - (n - 200 + 10) / (10 * 4)) ** 3
        assert float(np.float32(0.5 * (1 - ratio))) == ref[n]
    assert not PruneLayer().initted
    assert sorted(PruneLayer().state_dict().keys()) == ["_cur_sparsity", "_n_updates", "callback.t", "mask"]


def test_host_mirror_resyncs_after_external_writes():
    p = nn.Parameter(torch.zeros(1, dtype=torch.int), requires_grad=False)
    m = HostMirror()
    assert m.get(p) == 0
    p += 1
    m.wrote(p, 1)
    assert m.get(p) == 1
    with torch.no_grad():
        p.copy_(torch.tensor([41], dtype=torch.int))     # load_state_dict-style in-place write
    assert m.get(p) == 41
    p2 = nn.Parameter(torch.tensor([7], dtype=torch.int), requires_grad=False)   # preload-style replacement
    assert m.get(p2) == 7


def test_quantize_and_prune_factories():
    conv = qs.quantize(qs.prune(nn.Conv2d(4, 6, 3)), bits=8, bias_bits=8, channelwise=0)
    assert conv.__class__.__name__ == "Conv2d"
    assert conv.quantize.channelwise == 0 and conv.quantize_bias.channelwise == 0
    assert conv.quantize.callback is conv.quantize_bias.callback          # shared instance
    assert conv.quantize.batch_dimension == -1
    assert "prune.mask" in conv.state_dict() and "quantize_bias" not in conv.state_dict()
    with pytest.raises(ValueError):
        qs.quantize(torch.rand(10))
    with pytest.raises(ValueError):
        qs.prune(torch.rand(10))
    layer = qs.quantize(bits=4, timeout=0)
    assert str(layer) == "QuantizeLayer(bits=4, timeout=0, callback=ScalerQuantizer, channelwise=1)"
    assert str(qs.prune()) == "PruneLayer(sparsity=0.5, start=1000, interval=1000, repetition=4, dimensions={1})"


class LeNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 6, 5)
        self.conv2 = nn.Conv2d(6, 16, 5)
        self.fc = nn.Sequential(nn.Linear(400, 120), nn.ReLU(), nn.Linear(120, 10))
        self.act = nn.ReLU()


def test_convert_structure_and_filters():
    net = LeNet()
    with contextlib.redirect_stdout(io.StringIO()):
        p = qs.convert(net, qs.prune(sparsity=0.5), activation_layers=[nn.ReLU], inplace=False)
        q = qs.convert(p, qs.quantize(bits=8, timeout=10), weight_layers=[nn.Conv2d, nn.Linear],
                       activation_layers=[nn.ReLU], excluded_weight_layer_indexes=[(nn.Linear, [-1])], input=True,
                       inplace=False)
    assert isinstance(net.act, nn.ReLU)                       # inplace=False leaves the original alone
    inner = q[1]
    # activation: Sequential(Sequential(ReLU, PruneLayer), QuantizeLayer)  (qsparse/convert.py:214-217)
    assert isinstance(inner.act, nn.Sequential) and isinstance(inner.act[1], qs.quantize.__globals__["QuantizeLayer"])
    assert isinstance(inner.act[0][1], qs.prune.__globals__["PruneLayer"])
    assert hasattr(inner.conv1, "quantize") and hasattr(inner.conv2, "quantize")
    assert hasattr(inner.fc[0], "quantize") and not hasattr(inner.fc[2], "quantize")     # last Linear excluded
    names = [n for n, m in q.named_modules() if n.endswith("prune") or n.endswith("quantize")]
    assert "1.conv1.quantize" in names
    assert inner.conv1.quantize.name == "1.conv1.quantize"    # auto naming
    # callbacks are deep-copied per layer by convert (qsparse/convert.py:109-116)
    assert inner.conv1.quantize.callback is not inner.conv2.quantize.callback
    with contextlib.redirect_stdout(io.StringIO()):
        r = qs.convert(LeNet(), qs.quantize(bits=8), weight_layers=[nn.Conv2d], include="conv2")
    assert hasattr(r.conv2, "quantize") and not hasattr(r.conv1, "quantize")
    with contextlib.redirect_stdout(io.StringIO()):
        r = qs.convert(LeNet(), qs.quantize(bits=8), activation_layers=[nn.ReLU], order="pre", exclude="fc")
    assert isinstance(r.act[0], qs.quantize.__globals__["QuantizeLayer"]) and isinstance(r.fc[1], nn.ReLU)
    with pytest.raises(AssertionError):
        qs.convert(LeNet(), nn.ReLU())


def test_devise_layerwise_schedule_keeps_reference_behaviour():
    with contextlib.redirect_stdout(io.StringIO()):
        net = qs.convert(LeNet(), qs.prune(sparsity=0.5), activation_layers=[nn.ReLU])
        net = qs.devise_layerwise_pruning_schedule(net, start=10, interval=100, mask_refresh_interval=10)
    from qsparse_b200.sparse import PruneLayer
    layers = [m for m in net.modules() if isinstance(m, PruneLayer)]
    assert [l.start for l in layers] == sorted(l.start for l in layers) == [10, 111]
    assert all(l.repetition == 1 and l.schedules == [l.start] for l in layers)
    assert all(l.rampup_interval == 1000 for l in layers)     # left untouched, as in the reference


def test_unpack_int4_reader_side():
    """unpack_int4 inverts the nibble packing of qsb_quant_export_int4 (low nibble = even element; kinds decimal /
    scaler are 4-bit two's complement, line is unsigned), for even and odd element counts."""
    from qsparse_b200.quantize import unpack_int4
    rng = np.random.default_rng(0)
    for shape in ((3, 5, 7), (4, 6)):
        n = int(np.prod(shape))
        for kind, lo, hi in (("decimal", -8, 8), ("scaler", -8, 8), ("line", 0, 16)):
            codes = rng.integers(lo, hi, n)
            nib = np.concatenate([codes & 0xF, [0] * (n % 2)]).astype(np.uint8)
            packed = torch.from_numpy((nib[0::2] | (nib[1::2] << 4)).astype(np.uint8))
            got = unpack_int4(dict(q=packed, kind=kind, shape=shape))
            assert got.dtype == (torch.uint8 if kind == "line" else torch.int8)
            assert np.array_equal(got.numpy().reshape(-1), codes)


def test_uniform_callback_device_rng_flag():
    import qsparse_b200 as q
    assert q.UniformPruningCallback().device_rng is False
    cb = q.UniformPruningCallback(mask_refresh_interval=3, device_rng=True)
    assert cb.device_rng is True and cb.mask_refresh_interval == 3


def test_layerwise_schedule_reference_behaviour_and_fixed_variant():
    """devise_layerwise_pruning_schedule replicates the reference (rampup_interval untouched: the ramp formula
    overshoots); fix_ramp=True makes the single schedule point reach exactly `sparsity`."""
    def net():
        return nn.Sequential(qs.prune(sparsity=0.5, start=100, interval=100, repetition=4),
                             qs.prune(sparsity=0.75, start=100, interval=100, repetition=4))

    def target(layer):
        n = layer.schedules[0]
        ratio = (1.0 - (n - layer.start + layer.rampup_interval) / (layer.interval * layer.repetition)) ** 3
        return layer.sparsity * (1 - ratio)

    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        kept = qs.devise_layerwise_pruning_schedule(net(), start=1, interval=10)
        fixed = qs.devise_layerwise_pruning_schedule(net(), start=1, interval=10, fix_ramp=True)
    assert [l.start for l in kept] == [1, 12] and [l.schedules for l in kept] == [[1], [12]]
    assert [l.rampup_interval for l in kept] == [100, 100]            # untouched, like the reference
    assert target(kept[0]) == 0.5 * (1 - (1.0 - 100 / 10) ** 3)       # = 365: the reference's overshoot
    assert [l.rampup_interval for l in fixed] == [10, 10]
    assert target(fixed[0]) == 0.5 and target(fixed[1]) == 0.75
    for a, b in zip(kept, fixed):
        assert (a.start, a.interval, a.repetition, a.schedules) == (b.start, b.interval, b.repetition, b.schedules)


def test_graph_mode_host_logic():
    """qsparse_b200.graphs without a GPU: the mode flag nests and restores, routes that take a host step index refuse
    to run in graph mode, the quantizers' call count is read the same way for the int and the Parameter form."""
    import torch
    from qsparse_b200 import graphs
    from qsparse_b200.quantize import AdaptiveQuantizer, DecimalQuantizer
    assert not graphs.active()
    graphs.require_eager("anything")                     # a no-op outside graph mode
    with graphs.graph_mode():
        assert graphs.active()
        with graphs.graph_mode():
            assert graphs.active()
        assert graphs.active()
        with pytest.raises(graphs.NotCapturable):
            graphs.require_eager("a route with a host step index")
    assert not graphs.active()
    with pytest.raises(ZeroDivisionError):
        with graphs.graph_mode():
            1 / 0
    assert not graphs.active()                           # restored on the way out of an exception too
    assert issubclass(graphs.NotCapturable, RuntimeError)
    d = DecimalQuantizer()
    d.t = 7
    assert graphs.host_index(d) == 7
    a = AdaptiveQuantizer()
    a.t = torch.nn.Parameter(torch.zeros(1), requires_grad=False)   # what its first optimize() creates (ref :423-425)
    a._t_host = 3
    assert graphs.host_index(a) == 3
    assert graphs.NO_REFRESH_INTERVAL == 1 << 30
    import qsparse_b200
    assert qsparse_b200.GraphedTrainStep is graphs.GraphedTrainStep
