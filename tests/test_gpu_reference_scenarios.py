"""The behaviours the reference's own test-suite pins (tests/test_quantize.py:15-148,
tests/test_sparse.py:43-202, tests/test_util.py:15-111, tests/test_convert.py:34-145), re-stated against
``qsparse_b200`` on CUDA tensors: the same user-level scenarios (layer schedules, weight / bias wrapping,
integer-arithmetic equivalence, adaptive + group-wise quantization, structured / unstructured / uniform /
gradient / l0 pruning, conversion, naming, options, checkpoint preload), written from scratch."""
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rand(*shape, seed=0, lo=0.0, hi=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.rand(*shape, device=DEV, generator=g) * (hi - lo) + lo


def zero_fraction(t) -> float:
    return 1.0 - torch.count_nonzero(t).item() / t.numel()


def scale_to_decimal(scale):
    return (1 / scale).nan_to_num(posinf=1, neginf=1).log2().round()


class SmallLeNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv2d(3, 6, 5), nn.Conv2d(6, 16, 5)
        self.fc1, self.fc2, self.fc3 = nn.Linear(400, 120), nn.Linear(120, 84), nn.Linear(84, 10)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.conv1(x)), 2)
        x = F.max_pool2d(F.relu(self.conv2(x)), 2)
        x = F.relu(self.fc1(x.flatten(1)))
        return self.fc3(F.relu(self.fc2(x)))


# ----------------------------------------------------------------------------- quantize
def test_activation_layer_equals_functional_after_timeout():
    import qsparse_b200 as q
    from qsparse_b200.quantize import quantize_with_decimal
    x = rand(1, 10, 32, 32, seed=1, lo=-2, hi=2)
    layer = q.quantize(bits=8, timeout=5, channelwise=-1, callback=q.DecimalQuantizer())
    for _ in range(6):
        y = layer(x)
    want = quantize_with_decimal(x, bits=8, decimal=scale_to_decimal(layer.weight), channel_index=-1)
    assert torch.equal(y, want)
    assert not torch.equal(y, x)


def test_weight_and_bias_wrapping_schedule_and_errors():
    import qsparse_b200 as q
    from qsparse_b200.quantize import quantize_with_scaler
    x = rand(1, 10, 32, 32, seed=2)

    def make():
        torch.manual_seed(0)
        return q.quantize(nn.Conv2d(10, 30, 3).to(DEV), bits=8, bias_bits=8, timeout=5, callback=q.ScalerQuantizer(),
                          channelwise=0)
    conv = make()
    conv.train()
    for _ in range(6):
        conv(x)
    raw_w, raw_b = conv._parameters["weight"], conv._parameters["bias"]
    assert torch.equal(conv.weight, quantize_with_scaler(raw_w, 8, conv.quantize.weight, channel_index=0))
    assert torch.equal(conv.bias, quantize_with_scaler(raw_b, 8, conv.quantize_bias.weight, channel_index=0))
    assert not torch.equal(raw_w, conv.weight)            # the Parameter keeps full precision
    # in eval mode the schedule never fires
    conv = make()
    conv.eval()
    for _ in range(12):
        conv(x)
    assert torch.equal(conv.weight, conv._parameters["weight"])
    with pytest.raises(ValueError):
        q.quantize(torch.rand(10))


def test_integer_arithmetic_equivalence():
    """8-bit pow2 weights x 8-bit pow2 inputs: the float pipeline equals pure integer arithmetic."""
    import qsparse_b200 as q
    from qsparse_b200.quantize import quantize_with_decimal
    ni, no = 7, 6
    g = torch.Generator(device=DEV).manual_seed(3)
    xi = torch.randint(-128, 127, (3, 10, 32, 32), device=DEV, generator=g)
    xf = xi.float() / 2 ** ni
    torch.manual_seed(1)
    conv = q.quantize(nn.Conv2d(10, 30, 3, bias=False).to(DEV), bits=8, timeout=5, channelwise=0,
                      callback=q.DecimalQuantizer())
    conv.train()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False               # exact fp32 accumulation of small integers
    try:
        for _ in range(6):
            conv(xf)
        y_float = quantize_with_decimal(conv(xf), 8, no)
        dec = scale_to_decimal(conv.quantize.weight).int()
        w_int = (conv.weight * (2.0 ** dec).view(-1, 1, 1, 1)).round().long()
        acc = F.conv2d(xi.double(), w_int.double()).long()    # integer convolution (exact in fp64)
        for c in range(acc.shape[1]):
            acc[:, c] = (acc[:, c].double() / 2.0 ** (ni + int(dec[c]) - no)).long()
        assert torch.equal(y_float.double(), acc.double() / 2 ** no)
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_adaptive_quantizer_train_and_eval_forms():
    import qsparse_b200 as q
    from qsparse_b200.quantize import quantize_with_line
    x = rand(1, 10, 32, 32, seed=4, lo=-2, hi=2)
    layer = q.quantize(bits=8, timeout=5, channelwise=1, callback=q.AdaptiveQuantizer())
    for _ in range(6):
        y = layer(x)
    assert torch.equal(y, quantize_with_line(x, bits=8, lines=layer.weight, channel_index=1))
    layer.eval()
    assert torch.equal(layer(x), quantize_with_line(x, bits=8, lines=layer.weight, channel_index=1,
                                                    float_zero_point=False))


def test_groupwise_quantization():
    import qsparse_b200 as q
    from qsparse_b200.quantize import quantize_with_line
    x = rand(16, 10, 16, 16, seed=5, lo=-2, hi=2) * torch.linspace(0.2, 3, 10, device=DEV).view(1, 10, 1, 1)
    layer = q.quantize(bits=8, timeout=5, channelwise=1, callback=q.AdaptiveQuantizer(group_num=4, group_timeout=20))
    for _ in range(35):
        y = layer(x)
    from qsparse_b200 import ops
    from tests.conftest import ulp_diff
    grouped = layer.weight.clone()
    for gi in range(4):                      # the reference's loop (quantize.py:361-366), fp32 torch mean
        member = layer.callback.groups == gi
        grouped[member] = grouped[member].mean(dim=0)
    # the layer shares the group means through ONE device kernel that rounds the exact mean once: within 1 ulp of
    # torch's fp32 mean (whose own summation order differs between CPU and CUDA), non-pow2 parameter tolerance
    mine = ops.group_mean(layer.weight.data, layer.callback.groups, 4)
    assert ulp_diff(mine.cpu().numpy(), grouped.cpu().numpy()).max() <= 1
    assert torch.equal(y, quantize_with_line(x, bits=8, lines=mine, channel_index=1))
    assert len(torch.unique(mine, dim=0)) <= 4


# ----------------------------------------------------------------------------- prune
def test_activation_pruning_structured_uniform_and_shape_rules():
    import qsparse_b200 as q
    start, interval, rep = 5, 2, 3
    x, x2 = rand(1, 10, 32, 32, seed=6), rand(1, 10, 64, 64, seed=7)
    layer = q.prune(sparsity=0.5, start=start, interval=interval, repetition=rep)
    for _ in range(start + interval * (rep + 1)):
        y = layer(x)
    assert abs(zero_fraction(y) - 0.5) <= 1 / y.numel()
    assert torch.equal(y == 0, (layer.mask == 0).expand_as(y))
    # unstructured uniform pruning; the last call in eval mode
    np.random.seed(0)
    layer = q.prune(sparsity=0.5, start=start, interval=interval, repetition=rep, dimensions={0, 1, 2, 3},
                    callback=q.UniformPruningCallback())
    layer.train()
    steps = start + interval * (rep + 2)
    for i in range(steps):
        if i == steps - 1:
            layer.eval()
        y = layer(x)
    assert abs(zero_fraction(y) - 0.5) <= 1 / y.numel()
    assert torch.equal(y == 0, layer.mask == 0)
    with pytest.raises(RuntimeError):
        layer(x2)                                         # an unstructured mask fixes the input shape
    # the device-RNG extension keeps the invariants: exact budget per refresh, pruned positions stay pruned
    torch.manual_seed(4)
    layer = q.prune(sparsity=0.5, start=start, interval=interval, repetition=rep, dimensions={0, 1, 2, 3},
                    callback=q.UniformPruningCallback(device_rng=True))
    layer.train()
    prev_dead = torch.zeros_like(x, dtype=torch.bool)
    for i in range(steps):
        y = layer(x)
        if i >= start:
            dead = layer.mask == 0
            assert torch.equal(dead & prev_dead, prev_dead)
            prev_dead = dead
    assert abs(zero_fraction(y) - 0.5) <= 1 / y.numel() and torch.equal(y == 0, layer.mask == 0)
    # a channel mask lets the spatial size change in eval mode
    layer = q.prune(sparsity=0.5, start=start, interval=interval, repetition=rep, dimensions={1})
    for _ in range(start + interval * (rep + 1)):
        layer(x)
    layer.eval()
    y2 = layer(x2)
    assert y2.shape == x2.shape and abs(zero_fraction(y2) - 0.5) <= 4 / y2.numel()


def test_weight_pruning_variants():
    import qsparse_b200 as q
    start, interval, rep = 5, 2, 3
    x = rand(1, 10, 32, 32, seed=8)
    steps = start + interval * (rep + 1)

    def conv():
        torch.manual_seed(2)
        return nn.Conv2d(10, 30, 3).to(DEV)
    pc = q.prune(conv(), sparsity=0.5, start=start, interval=interval, repetition=rep,
                 callback=q.MagnitudePruningCallback(running_average=False))
    pc.train()
    for _ in range(steps):
        pc(x)
    assert abs(zero_fraction(pc.weight) - 0.5) <= 1 / pc.weight.numel()
    assert torch.equal(pc.weight == 0, (pc.prune.mask == 0).expand_as(pc.weight))
    assert zero_fraction(dict(pc.named_parameters())["weight"]) < 0.4        # the raw weight is untouched
    np.random.seed(1)
    pc = q.prune(conv(), sparsity=0.5, start=start, interval=interval, repetition=rep,
                 callback=q.UniformPruningCallback())
    pc.train()
    for _ in range(steps):
        pc(x)
    assert abs(zero_fraction(pc.weight) - 0.5) <= 1 / pc.weight.numel()
    pc.eval()
    assert torch.equal(pc.weight == 0, (pc.prune.mask == 0).expand_as(pc.weight))
    pc = q.prune(conv(), sparsity=0.5, start=start, interval=interval, repetition=rep)
    pc.eval()
    for _ in range(steps):
        pc(x)
    assert zero_fraction(pc.weight) < 0.4                 # the schedule only runs while training
    with pytest.raises(ValueError):
        q.prune(torch.rand(10))


def test_gradient_and_l0_importance():
    import qsparse_b200 as q
    shape = (3, 24, 24)
    mean = rand(*shape, seed=9)
    mask = torch.ones((1,) + shape, dtype=torch.bool, device=DEV)
    cb = q.MagnitudePruningCallback(use_gradient=True)
    g = torch.Generator(device=DEV).manual_seed(10)
    for _ in range(200):
        inp = (mean + torch.randn(shape, device=DEV, generator=g)).view(1, *shape).requires_grad_(True)
        out = cb(inp, 0.5, mask)
        out.backward(torch.rand((1,) + shape, device=DEV, generator=g) / 10)
    assert abs(zero_fraction(mask) - 0.5) <= 2 / mask.numel()
    mask = torch.ones(shape, dtype=torch.bool, device=DEV)
    cb = q.MagnitudePruningCallback(l0=True)
    for _ in range(1500):      # long enough for the running averages of the 0/1 indicators to separate
        cb((torch.rand(shape, device=DEV, generator=g) > 0.5).float(), 0.5, mask)
    assert abs(zero_fraction(mask) - 0.5) <= 2 / mask.numel()


def test_layerwise_schedule_is_monotone():
    import qsparse_b200 as q
    from qsparse_b200.sparse import PruneLayer
    net = q.convert(SmallLeNet(), q.prune(sparsity=0.5, callback=q.MagnitudePruningCallback()),
                    activation_layers=[nn.Conv2d, nn.Linear], log=False)
    net = q.devise_layerwise_pruning_schedule(net, start=10, interval=100, mask_refresh_interval=10)
    starts = [m.start for m in net.modules() if isinstance(m, PruneLayer)]
    assert len(starts) == 5 and starts == sorted(starts) and len(set(starts)) == 5


# ----------------------------------------------------------------------------- util / convert
def test_auto_naming_and_options(capsys):
    import qsparse_b200 as q

    class Two(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear1 = q.quantize(q.prune(nn.Linear(10, 30)))
            self.linear2 = q.quantize(q.prune(nn.Linear(30, 1)))
    net = q.auto_name_prune_quantize_layers(Two())
    assert [net.linear1.prune.name, net.linear1.quantize.name, net.linear2.prune.name, net.linear2.quantize.name] == \
        ["linear1.prune", "linear1.quantize", "linear2.prune", "linear2.quantize"]
    try:
        q.set_qsparse_options()
        q.set_qsparse_options(log_on_created=False)
        assert q.get_qsparse_option("log_on_created") is False
        capsys.readouterr()
        q.prune(sparsity=0.5)
        q.quantize(bits=8)
        seen = capsys.readouterr()
        assert "[Prune" not in seen.out + seen.err and "[Quantize" not in seen.out + seen.err
        q.set_qsparse_options(log_during_train=False)
        assert q.get_qsparse_option("log_during_train") is False
    finally:
        q.set_qsparse_options(log_on_created=True, log_during_train=True)


def test_squeeze_and_mask_helpers():
    import qsparse_b200 as q
    from qsparse_b200.util import squeeze_tensor_to_shape
    t = rand(10, 30, 7, 8, seed=11)
    assert tuple(squeeze_tensor_to_shape(t, (1, 30, 7, 1)).shape) == (1, 30, 7, 1)
    mask = q.calculate_mask_given_importance(t, 0.47)
    assert 1 - mask.sum().item() / mask.numel() == 0.47


def test_preload_state_dict_then_load():
    import qsparse_b200 as q
    from qsparse_b200.util import preload_qsparse_state_dict

    def make():
        torch.manual_seed(4)
        return q.quantize(q.prune(nn.Conv2d(16, 32, 3).to(DEV), sparsity=0.5, start=20, interval=5, repetition=4),
                          bits=8, timeout=10)
    conv = make()
    for i in range(45):
        conv(rand(10, 16, 7, 7, seed=100 + i))
    with pytest.raises(RuntimeError):
        make().load_state_dict(conv.state_dict())         # lazily shaped parameters do not match yet
    other = make()
    preload_qsparse_state_dict(other, conv.state_dict())
    other.load_state_dict(conv.state_dict())
    a, b = conv.state_dict(), other.state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k].cpu(), b[k].cpu()) for k in a)
    assert other.prune._n_mirror.get(other.prune._n_updates) == 45      # host mirrors re-read the loaded counters
    other(rand(10, 16, 7, 7, seed=999))                                  # and the layer keeps running
    assert other.prune._n_updates.item() == 46


def test_convert_scenarios():
    import qsparse_b200 as q
    skip = [(nn.Conv2d, [0]), (nn.Linear, [-1])]
    pruned = q.convert(SmallLeNet(), q.prune(sparsity=0.5, callback=q.MagnitudePruningCallback()),
                       weight_layers=[nn.Conv2d, nn.Linear], activation_layers=[nn.Conv2d, nn.Linear],
                       excluded_weight_layer_indexes=skip, excluded_activation_layer_indexes=skip, log=False)
    both = q.convert(pruned, q.quantize(bits=8), weight_layers=[nn.Conv2d, nn.Linear],
                     activation_layers=[nn.Conv2d, nn.Linear], input=True, log=False)
    mods = dict(both.named_modules())
    assert [type(mods[k]).__name__ for k in ("1.fc1.0.0.prune", "1.fc1.0.0.quantize", "1.fc1.0.1", "1.fc1.1")] == \
        ["PruneLayer", "QuantizeLayer", "PruneLayer", "QuantizeLayer"]
    # nothing requested -> nothing changes (and a warning)
    plain = SmallLeNet()
    with pytest.warns(UserWarning):
        assert str(q.convert(plain, q.prune(sparsity=0.5), log=False)) == str(SmallLeNet())
    # DataParallel-style wrapper
    wrapped = q.convert(nn.DataParallel(SmallLeNet()), q.quantize(bits=8), weight_layers=[nn.Conv2d, nn.Linear],
                        activation_layers=[nn.Conv2d, nn.Linear], log=False)
    assert "quantize" in str(wrapped).lower()
    # include filter
    net = nn.Sequential(OrderedDict(conv1=nn.Conv2d(3, 6, 5), special=nn.Sequential(nn.Conv2d(6, 16, 5))))
    text = str(q.convert(net, q.quantize(bits=8), weight_layers=[nn.Conv2d], include=["special"], log=False))
    assert text.count("quantize") == 1 and text.index("special") < text.index("quantize")
    # order="pre" puts the operator in front; a second conversion nests around the first
    net = nn.Sequential(OrderedDict(conv1=nn.Conv2d(3, 6, 5), fc1=nn.Linear(84, 10)))
    pre = q.convert(net, q.quantize(bits=8), activation_layers=[nn.Conv2d, nn.Linear], order="pre", log=False)
    text = str(pre).lower()
    assert text.count("quantizelayer") == 2 and text.index("quantize") < text.index("conv2d")
    tail = text[text.index("conv2d"):]
    assert tail.index("linear") > tail.index("quantize")
    nested = str(q.convert(pre, q.prune(sparsity=0.5), activation_layers=[nn.Conv2d, nn.Linear], log=False)).lower()
    assert nested.index("quantize") < nested.index("prune")


def test_adaptive_batched_channel0_raises_like_the_reference():
    """AdaptiveQuantizer.optimize(batched=True, channel_index=0): the reference's transpose(1, 0).view(...) raises a
    RuntimeError for every input (quantize.py:399-402); so does this implementation."""
    import qsparse_b200 as q
    with pytest.raises(RuntimeError, match="view size is not compatible"):
        q.AdaptiveQuantizer().optimize(rand(4, 3, 5, 5, seed=1), 8, None, channel_index=0, batched=True)
