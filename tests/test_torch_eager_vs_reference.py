"""oracle/torch_eager.py (the reference's ATen op sequence restated as plain torch ops — the
"PyTorch-eager on the same B200" baseline of bench.py) against the imported reference itself, on CPU
tensors.  Runs only where /root/reference exists (the build container)."""
import sys
from pathlib import Path

import pytest
import torch

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "qsparse").exists(), reason="reference tree not present")


def _ref():
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    import qsparse
    from qsparse.quantize import AdaptiveQuantizer, DecimalQuantizer
    from qsparse.sparse import MagnitudePruningCallback
    qsparse.set_qsparse_options(log_on_created=False)
    return qsparse, DecimalQuantizer, AdaptiveQuantizer, MagnitudePruningCallback


def test_prune_quantize_step_matches_reference_layers():
    from oracle.torch_eager import PruneQuantizeEager
    qsparse, DecimalQuantizer, _, MagnitudePruningCallback = _ref()
    torch.manual_seed(0)
    C = 12
    cb = MagnitudePruningCallback()
    cb.train()
    mask = torch.ones(1, C, 1, 1, dtype=torch.bool)
    q = qsparse.quantize(bits=8, channelwise=-1, timeout=1, callback=DecimalQuantizer())
    q.train()
    q(torch.zeros(2, C, 3, 3))      # step 0 of the quantize layer is the pass-through before its timeout
    eager = PruneQuantizeEager(C, "cpu", sparsity=0.75, bits=8)
    scale = torch.linspace(0.2, 2.0, C).view(1, C, 1, 1)
    for t in range(5):
        x = torch.relu(torch.randn(6, C, 9, 9)) * scale
        g = torch.randn(6, C, 9, 9) * 3
        xr = x.clone().requires_grad_(True)
        yr = q(cb(xr, 0.75, mask))
        yr.backward(g.clone())
        ye = eager.forward(x.clone())
        gxe = eager.backward(g.clone())
        assert torch.equal(yr.detach(), ye), t
        assert torch.equal(xr.grad, gxe), t
        assert torch.equal(mask, eager.mask), t
        assert torch.equal(cb.magnitude, eager.magnitude), t
        assert torch.equal(q.weight.view(-1), eager.weight.view(-1)), t


def test_weight_line_quant_matches_reference_layer():
    from oracle.torch_eager import WeightLineQuantEager
    qsparse, _, AdaptiveQuantizer, _ = _ref()
    torch.manual_seed(1)
    ql = qsparse.quantize(torch.nn.Linear(64, 48), bits=4, channelwise=0, timeout=1, callback=AdaptiveQuantizer())
    ql.train()
    raw = dict(ql.named_parameters())["weight"]
    _ = ql.weight                   # pass-through access before the timeout
    eager = WeightLineQuantEager(bits=4)
    for t in range(3):
        with torch.no_grad():
            raw.copy_(torch.randn(48, 64) * 0.02 * (1 + 0.2 * t))
        assert torch.equal(ql.weight.detach(), eager.forward(raw.detach())), t
    assert torch.equal(ql.quantize.weight, eager.lines)


def test_unstructured_prune_matches_reference_callback():
    from oracle.torch_eager import UnstructuredPruneEager
    _, _, _, MagnitudePruningCallback = _ref()
    torch.manual_seed(2)
    w = torch.randn(40, 30, 3, 3) * 0.02
    cb = MagnitudePruningCallback(running_average=True)
    cb.train()
    mask = torch.ones(w.shape, dtype=torch.bool)
    eager = UnstructuredPruneEager(w, 0.5)
    for t in range(3):
        wt = w * (1 + 0.1 * t)
        yr = cb(wt, 0.5, mask)
        ye = eager.forward(wt)
        assert torch.equal(yr, ye) and torch.equal(mask, eager.mask), t


def test_mnist_eager_net_matches_reference_converted_net():
    """build_mnist_eager (the config-1 GPU-eager baseline of bench.py) trains step for step like the
    reference's convert()'ed net on CPU: identical losses for 35 steps (warm-up, quantization start at 10,
    pruning start at 20 and two ramp points)."""
    import torch.nn.functional as F
    from oracle.gen_golden_config1 import build, data
    from oracle.torch_eager import build_mnist_eager
    qsparse, DecimalQuantizer, _, _ = _ref()
    from qsparse.quantize import ScalerQuantizer
    for kind, factory in (("scaler", ScalerQuantizer), ("decimal", DecimalQuantizer)):
        ref = build(qsparse, factory)
        mine = build_mnist_eager(kind)
        ref.train(), mine.train()
        o1 = torch.optim.Adadelta(ref.parameters(), lr=1.0)
        o2 = torch.optim.Adadelta([p for p in mine.parameters() if p.requires_grad], lr=1.0)
        for step in range(35):
            x, y = data(step)
            o1.zero_grad(), o2.zero_grad()
            l1 = F.nll_loss(ref(x), y)
            l2 = F.nll_loss(mine(x), y)
            l1.backward(), l2.backward()
            o1.step(), o2.step()
            assert l1.item() == l2.item(), (kind, step, l1.item(), l2.item())
