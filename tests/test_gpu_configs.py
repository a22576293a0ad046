"""GPU parity at the level of BASELINE.json's configs: config 1 (MNIST net end to end against
goldens recorded from the reference), config 3 (4-bit channel-wise AdaptiveQuantizer on a
[4096,4096] weight) and config 4 (unstructured magnitude prune of a conv weight set) against
the CPU oracle."""
import contextlib
import importlib
import io
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import oracle as orc
from tests.conftest import bits_equal, ulp_diff

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def npy(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("tag,fuse", [("scaler", False), ("decimal", False), ("scaler", True), ("decimal", True)])
def test_config1_mnist_end_to_end(tag, fuse):
    """BASELINE config 1: 60 training steps of the converted MNIST net on the GPU (``fuse``: with the fusion
    pass applied — the steady-state steps of the two prune->quantize activation sites then run through the
    fused kernels, with the DEFAULT ScalerQuantizer callback as well as with DecimalQuantizer).

    (a) Teacher-forced parity: EVERY Prune/Quantize layer call of the run (2 + 8 layers x 60 steps)
        is checked against the oracle's restatement of the reference's layer logic on the identical
        input tensor: outputs and masks bit-exact, scales bit-exact, magnitudes <= 8 ulp.
    (b) Against the reference's own CPU run (tests/golden/config1_mnist.npz): schedules, step
        counters and the pruned fraction per step are identical, losses agree to 1e-4 until
        quantization starts; afterwards cuDNN-vs-CPU convolution noise is amplified by the quantized
        training (the channel magnitudes of this BatchNorm'ed net are tied to within 0.1-2 %), so
        scales / losses are compared with a tolerance and masks only on channels that are not near
        the threshold."""
    import qsparse_b200 as qs
    from oracle.gen_golden_config1 import run, build
    from tests.oracle_layers import OraclePrune, OracleQuantize
    q = importlib.import_module("qsparse_b200.quantize")
    sp = importlib.import_module("qsparse_b200.sparse")
    qs.set_qsparse_options(log_on_created=False)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # cuDNN's default weight-gradient kernels accumulate with atomics: the run-to-run noise is amplified by the
    # quantized training like the GPU-vs-CPU noise of (b) and made this test fail about once in twenty runs
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    ref = np.load(ROOT / "tests" / "golden" / "config1_mnist.npz")
    factory = {"scaler": q.ScalerQuantizer, "decimal": q.DecimalQuantizer}[tag]

    # ---- (a) teacher forcing: hook every layer call -------------------------------------------
    checked = {"prune": 0, "quant": 0}
    emulators = {}

    def hook(module, inputs, output):
        x = npy(inputs[0])
        if isinstance(module, sp.PruneLayer):
            em = emulators.setdefault(id(module), OraclePrune(0.5, 20, 10, 4))
            exp = em.forward(x)
            assert bits_equal(npy(output), exp), ("prune", module.name, em.n)
            if em.mag is not None:
                assert np.array_equal(npy(module.mask), em.mask), ("mask", module.name, em.n)
                assert ulp_diff(npy(module.callback.magnitude), em.mag).max() <= 8, ("mag", module.name, em.n)
            checked["prune"] += 1
        else:
            em = emulators.setdefault(id(module), OracleQuantize(8, 10, tag))
            exp = em.forward(x, module.training)
            assert bits_equal(npy(output), exp), ("quant", module.name, em.n)
            assert bits_equal(npy(module.weight).reshape(-1), em.weight), ("scale", module.name, em.n)
            checked["quant"] += 1

    from qsparse_b200.fused import FusedPruneQuantSequential
    captured, seen, fused_sites = {}, {}, []

    def capture(module, inputs, output):          # the activation in front of a fused site
        captured[id(module)] = npy(output)

    def site_hook(site, inputs, output):
        """a fused step bypasses the two layers' own forwards (and their hooks): check it here against the
        chained emulators — same state objects as the un-fused steps of the same layers use"""
        if site.fused_steps == seen.get(id(site), 0):
            return
        seen[id(site)] = site.fused_steps
        pre, p, qz = site._layers()
        x = captured[id(pre)]
        emp = emulators.setdefault(id(p), OraclePrune(0.5, 20, 10, 4))
        emq = emulators.setdefault(id(qz), OracleQuantize(8, 10, tag))
        exp = emq.forward(emp.forward(x), True)
        assert bits_equal(npy(output), exp), ("fused site", p.name, emp.n)
        assert np.array_equal(npy(p.mask), emp.mask), ("mask", p.name, emp.n)
        assert ulp_diff(npy(p.callback.magnitude), emp.mag).max() <= 8, ("mag", p.name, emp.n)
        assert bits_equal(npy(qz.weight).reshape(-1), emq.weight), ("scale", qz.name, emq.n)
        checked["prune"] += 1
        checked["quant"] += 1

    orig_build = build

    def build_hooked(qs_, cb):
        model = orig_build(qs_, cb)
        if fuse:
            qs_.fuse_prune_quantize(model)
        for m in model.modules():
            if isinstance(m, (sp.PruneLayer, q.QuantizeLayer)):
                m.register_forward_hook(hook)
            if isinstance(m, FusedPruneQuantSequential):
                fused_sites.append(m)
                m.register_forward_hook(site_hook)
                m._layers()[0].register_forward_hook(capture)
        return model

    import oracle.gen_golden_config1 as cfg1
    cfg1.build = build_hooked
    try:
        out, model = run(qs, factory, device="cuda")
    finally:
        cfg1.build = orig_build
    assert checked == {"prune": 2 * 60, "quant": 8 * 60}
    if fuse:   # steps 21-29, 31-39, 41-49, 51-59 of both sites are steady-state steps
        assert len(fused_sites) == 2 and all(s_.fused_steps == 36 for s_ in fused_sites), \
            [s_.fused_steps for s_ in fused_sites]

    # ---- (b) against the reference's CPU run --------------------------------------------------
    assert np.array_equal(out["sparsity"], ref[f"{tag}/sparsity"])
    for i in range(2):
        assert str(out[f"prune{i}_name"]) == str(ref[f"{tag}/prune{i}_name"])
        kr, ko = ref[f"{tag}/prune{i}_mask"].reshape(-1), out[f"prune{i}_mask"].reshape(-1)
        mr = ref[f"{tag}/prune{i}_magnitude"].reshape(-1)
        thr = np.sort(mr)[len(mr) // 2]
        decisive = np.abs(mr - thr) / thr > 0.08
        assert np.array_equal(kr[decisive], ko[decisive]), i
        assert (kr != ko).mean() <= 0.35
        assert np.allclose(out[f"prune{i}_magnitude"], ref[f"{tag}/prune{i}_magnitude"], rtol=0.1), i
    nq = len([k for k in ref.files if k.startswith(f"{tag}/quant") and k.endswith("_weight")])
    assert nq == 8      # input + 4 weights + 3 activations
    for i in range(nq):
        assert str(out[f"quant{i}_name"]) == str(ref[f"{tag}/quant{i}_name"])
        assert np.array_equal(out[f"quant{i}_n_updates"], ref[f"{tag}/quant{i}_n_updates"])
        assert np.allclose(out[f"quant{i}_weight"], ref[f"{tag}/quant{i}_weight"], rtol=0.1), i
    assert np.allclose(out["loss"][:6], ref[f"{tag}/loss"][:6], rtol=1e-3)        # first steps: only conv/BN noise
    assert np.allclose(out["loss"][:10], ref[f"{tag}/loss"][:10], rtol=5e-3)      # before quantization starts
    assert np.allclose(out["loss"], ref[f"{tag}/loss"], rtol=0.08)
    # state_dict layout (docs/advanced_usage.ipynb:848-856): keys / dtypes / shapes
    sd = model.state_dict()
    pk = [k for k in sd if k.endswith("callback.magnitude")][0].rsplit("callback.magnitude", 1)[0]
    assert sd[pk + "mask"].dtype == torch.bool and sd[pk + "_n_updates"].dtype == torch.int32
    assert sd[pk + "_cur_sparsity"].dtype == torch.float32 and sd[pk + "callback.t"].dtype == torch.int64
    assert sd[pk + "mask"].shape == sd[pk + "callback.magnitude"].shape


def test_config3_adaptive_channelwise_weight():
    """quantize(nn.Linear(4096,4096), bits=4, channelwise=0, AdaptiveQuantizer): three training
    accesses of .weight against the oracle (min/max reduce -> lines EMA -> line quant), bit-exact."""
    import qsparse_b200 as qs
    q = importlib.import_module("qsparse_b200.quantize")
    qs.set_qsparse_options(log_on_created=False)
    torch.manual_seed(3)
    lin = nn.Linear(4096, 4096)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(4096, 4096) * 0.02)
    w = lin.weight.detach().numpy().copy()
    lin = lin.cuda()
    with contextlib.redirect_stdout(io.StringIO()):
        ql = qs.quantize(lin, bits=4, channelwise=0, timeout=1, callback=q.AdaptiveQuantizer())
    ql.train()
    lines = np.zeros((4096, 2), np.float32)
    first = ql.weight                                   # step 0 < timeout: raw weight
    assert bits_equal(npy(first), w)
    for t in range(1, 4):
        got = ql.weight
        mn, mx = orc.minmax(w, 0)
        lines = orc.lines_ema(lines, mn, mx, t)
        assert bits_equal(npy(ql.quantize.weight), lines), t
        assert bits_equal(npy(got), orc.fq_line_fwd(w, lines, 4, 0, True)), t
    ql.eval()
    assert bits_equal(npy(ql.weight), orc.fq_line_fwd(w, lines, 4, 0, False))    # integer zero-point in eval
    assert len(np.unique(npy(ql.weight)[7])) <= 16                                 # 4 bits per row
    # gradient flows straight through (identity backward, quantize.py:183-185)
    ql.train()
    x = torch.randn(2, 4096, device="cuda")
    ql(x).sum().backward()
    assert ql._parameters["weight"].grad is not None


def test_config4_unstructured_weight_set():
    """Unstructured magnitude pruning (running-average magnitude, k-th value mask) of a set of conv
    weights, three refresh steps, against the oracle: masks and outputs bit-equal."""
    import qsparse_b200 as qs
    from qsparse_b200.sparse import MagnitudePruningCallback
    qs.set_qsparse_options(log_on_created=False)
    rng = np.random.default_rng(4)
    shapes = [(256, 128, 3, 3), (128, 128, 3, 3), (64, 32, 5, 5), (100, 37)]
    for shape in shapes:
        for sparsity in (0.5, 0.75):
            cb = MagnitudePruningCallback().cuda()
            cb.train()
            mask = nn.Parameter(torch.ones(*shape, dtype=torch.bool, device="cuda"), requires_grad=False)
            mag = np.zeros(shape, np.float32)
            ref_mask = np.ones(shape, bool)
            for t in range(4):
                w = (rng.standard_normal(shape) * 0.02).astype(np.float32)
                out = cb(torch.from_numpy(w).cuda(), sparsity, mask)
                mag = orc.magnitude_ema(mag, np.abs(w), t)
                if t > 0:
                    ref_mask, _ = orc.mask_given_importance(mag, sparsity)
                assert bits_equal(npy(cb.magnitude), mag), (shape, t)
                assert np.array_equal(npy(mask), ref_mask), (shape, t)
                assert bits_equal(npy(out), orc.mask_apply(w, ref_mask.reshape(-1))), (shape, t)
            pruned = 1 - npy(mask).mean()
            assert abs(pruned - sparsity) <= 1.0 / mask.numel() + 1e-12


def test_preload_state_dict_roundtrip():
    """reference tests/test_util.py:87-112 on CUDA tensors"""
    import qsparse_b200 as qs
    from qsparse_b200.util import preload_qsparse_state_dict
    qs.set_qsparse_options(log_on_created=False)

    def make_conv():
        with contextlib.redirect_stdout(io.StringIO()):
            return qs.quantize(qs.prune(nn.Conv2d(16, 32, 3), sparsity=0.5, start=200, interval=10, repetition=4),
                               bits=8, timeout=100).cuda()
    conv = make_conv()
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(241):
            conv(torch.rand(10, 16, 7, 7, device="cuda"))
    conv2 = make_conv()
    with pytest.raises(RuntimeError):
        conv2.load_state_dict(conv.state_dict())              # shapes unknown before the first forward
    conv3 = make_conv()
    preload_qsparse_state_dict(conv3, conv.state_dict())
    conv3.load_state_dict(conv.state_dict())
    x = torch.rand(10, 16, 7, 7, device="cuda")
    conv.eval()
    conv3.eval()
    conv3.quantize._quantized = True                          # not part of state_dict
    assert np.allclose(npy(conv(x)), npy(conv3(x)), atol=1e-5)
    assert int((~conv3.prune.mask).sum().item()) > 0


# ----------------------------------------------------------------------------- fusion pass
def _converted_net(fuse: bool, callback="decimal", via_convert=False):
    import qsparse_b200 as q
    torch.manual_seed(7)
    net = torch.nn.Sequential(
        torch.nn.Conv2d(3, 16, 3, padding=1), torch.nn.ReLU(),
        torch.nn.Conv2d(16, 24, 3, padding=1), torch.nn.ReLU(),
        torch.nn.Flatten(), torch.nn.Linear(24 * 8 * 8, 10)).cuda()
    net = q.convert(net, q.prune(sparsity=0.5, dimensions={1}, start=2, interval=3, repetition=2),
                    activation_layers=[torch.nn.ReLU], log=False)
    cb = {"decimal": q.DecimalQuantizer, "scaler": q.ScalerQuantizer, "default": lambda: None}[callback]()
    net = q.convert(net, q.quantize(bits=8, channelwise=-1, timeout=4, callback=cb),
                    activation_layers=[torch.nn.ReLU], log=False, fuse=fuse and via_convert)
    if fuse and not via_convert:
        net = q.fuse_prune_quantize(net)
    return net


@pytest.mark.parametrize("callback,via_convert", [("decimal", False), ("scaler", False), ("default", True)])
def test_fusion_pass_equals_unfused_layers(callback, via_convert):
    """fuse_prune_quantize keeps the module tree / state_dict keys and gives bit-identical outputs, gradients
    and layer state over warm-up, ramp, steady-state and eval steps; the fused route is actually taken."""
    from qsparse_b200.fused import FusedPruneQuantSequential
    torch.backends.cudnn.deterministic = True      # the two nets must see identical conv arithmetic
    torch.backends.cudnn.benchmark = False
    a, b = _converted_net(False, callback), _converted_net(True, callback, via_convert)
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    sites = [m for m in b.modules() if isinstance(m, FusedPruneQuantSequential)]
    assert len(sites) == 2
    a.train()
    b.train()
    for step in range(14):
        x = torch.randn(8, 3, 8, 8, device="cuda", generator=torch.Generator("cuda").manual_seed(100 + step))
        if step == 11:            # an eval step in between
            a.eval()
            b.eval()
        ya, yb = a(x), b(x)
        assert torch.equal(ya, yb), step
        if step != 11:
            ya.square().mean().backward()
            yb.square().mean().backward()
            for (na, pa), (nb, pb) in zip(a.named_parameters(), b.named_parameters()):
                if pa.grad is not None:
                    assert torch.equal(pa.grad, pb.grad), (step, na)
                    pa.grad = pb.grad = None
        sa, sb = a.state_dict(), b.state_dict()
        assert sa.keys() == sb.keys()
        for key in sa:
            assert torch.equal(sa[key], sb[key]), (step, key)
        a.train()
        b.train()
    assert all(s.fused_steps >= 6 for s in sites), [s.fused_steps for s in sites]
    # checkpoint interchange: the unfused model loads the fused model's state and vice versa
    a.load_state_dict(b.state_dict())
    b.load_state_dict(a.state_dict())


def test_weight_chain_fusion_equals_unfused():
    """quantize(prune(conv)) with an unstructured mask that freezes after `stop_mask_refresh`: the fusion pass
    sends the frozen-mask training steps through K8 with the element mask; outputs, weight gradients and all
    layer state stay bit-identical to the un-fused chain over warm-up, refresh, frozen and eval steps."""
    import qsparse_b200 as q
    q.set_qsparse_options(log_on_created=False)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False

    def make(cb_name, fuse):
        torch.manual_seed(11)
        conv = torch.nn.Conv2d(16, 32, 3, padding=1).cuda()
        cb = {"decimal": q.DecimalQuantizer, "scaler": q.ScalerQuantizer, "adaptive": q.AdaptiveQuantizer}[cb_name]()
        pcb = q.MagnitudePruningCallback(mask_refresh_interval=1, stop_mask_refresh=4)
        conv = q.prune(conv, sparsity=0.6, dimensions={0, 1, 2, 3}, start=1, interval=1, repetition=1, callback=pcb)
        conv = q.quantize(conv, bits=6, channelwise=0, timeout=2, callback=cb)
        net = torch.nn.Sequential(conv).cuda()
        if fuse:
            q.fuse_prune_quantize(net)
        return net

    for cb_name in ("decimal", "scaler", "adaptive"):
        a, b = make(cb_name, False), make(cb_name, True)
        a.train(), b.train()
        for step in range(12):
            x = torch.randn(4, 16, 10, 10, device="cuda", generator=torch.Generator("cuda").manual_seed(50 + step))
            if step == 9:
                a.eval(), b.eval()
            ya, yb = a(x), b(x)
            assert torch.equal(ya, yb), (cb_name, step)
            if step != 9:
                ya.square().mean().backward()
                yb.square().mean().backward()
                for (na, pa), (nb, pb) in zip(a.named_parameters(), b.named_parameters()):
                    if pa.grad is not None:
                        assert torch.equal(pa.grad, pb.grad), (cb_name, step, na)
                        # a tiny SGD step so that the weights (and with them the statistics) move
                        with torch.no_grad():
                            pa.sub_(0.05 * pa.grad)
                            pb.sub_(0.05 * pb.grad)
                        pa.grad = pb.grad = None
            sa, sb = a.state_dict(), b.state_dict()
            assert sa.keys() == sb.keys()
            for key in sa:
                assert torch.equal(sa[key], sb[key]), (cb_name, step, key)
            a.train(), b.train()
        assert b[0].fused_weight_steps >= 5, (cb_name, b[0].fused_weight_steps)


def test_weight_set_pruner_equals_per_layer_path():
    """WeightSetPruner.step() (one batched launch sequence for all pruned weights) followed by the forward ==
    the per-layer path, bit for bit: outputs, weight gradients, masks, magnitudes, counters — over warm-up, the
    ramp points (left to the layers themselves), steady state and an eval step."""
    import qsparse_b200 as q
    q.set_qsparse_options(log_on_created=False)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False

    def make():
        torch.manual_seed(3)
        layers = [torch.nn.Conv2d(8, 64, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(64, 64, 3, padding=1),
                  torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(64 * 6 * 6, 300)]
        net = torch.nn.Sequential(*layers).cuda()
        for i in (0, 2, 5):
            net[i] = q.prune(net[i], sparsity=0.5, dimensions={0, 1, 2, 3} if i != 5 else {0, 1}, start=2, interval=3,
                             repetition=2)
        return net

    a, b = make(), make()
    pruner = q.WeightSetPruner(b)
    assert len(pruner.layers) == 3
    a.train(), b.train()
    batched = 0
    for step in range(14):
        x = torch.randn(4, 8, 6, 6, device="cuda", generator=torch.Generator("cuda").manual_seed(200 + step))
        if step == 10:
            a.eval(), b.eval()
        batched += pruner.step()
        ya, yb = a(x), b(x)
        assert torch.equal(ya, yb), step
        if step != 10:
            ya.square().mean().backward()
            yb.square().mean().backward()
            for (na, pa), (nb, pb) in zip(a.named_parameters(), b.named_parameters()):
                if pa.grad is not None:
                    assert torch.equal(pa.grad, pb.grad), (step, na)
                    with torch.no_grad():
                        pa.sub_(0.05 * pa.grad)
                        pb.sub_(0.05 * pb.grad)
                    pa.grad = pb.grad = None
        sa, sb = a.state_dict(), b.state_dict()
        for key in sa:
            assert torch.equal(sa[key], sb[key]), (step, key)
        a.train(), b.train()
    assert batched >= 3 * 6, batched


def test_unstructured_callback_fused_step_equals_unfused():
    """MagnitudePruningCallback on a weight-shaped mask: the K9 route (one streaming pass) and the
    update_magnitude -> kth_value -> mask_build_apply route give identical magnitudes, masks, outputs and
    gradients, step after step (incl. the step-0 no-refresh rule and a non-contiguous input)."""
    import importlib
    sp = importlib.import_module("qsparse_b200.sparse")
    res = {}
    for fuse in (True, False):
        sp.FUSE_PRUNE_STEP = fuse
        try:
            g = torch.Generator(device="cuda").manual_seed(3)
            cb = sp.MagnitudePruningCallback(running_average=True)
            cb.train()
            mask = torch.nn.Parameter(torch.ones(96, 64, 3, 3, dtype=torch.bool, device="cuda"), requires_grad=False)
            rec = []
            for t in range(5):
                w = (torch.randn(96, 64, 3, 3, device="cuda", generator=g) * 0.05).requires_grad_(True)
                x = w.transpose(2, 3) if t == 3 else w          # a non-contiguous view once
                y = cb(x, 0.6, mask)
                y.sum().backward()
                rec.append((y.detach().clone(), w.grad.clone(), mask.data.clone(), cb.magnitude.data.clone()))
            res[fuse] = rec
        finally:
            sp.FUSE_PRUNE_STEP = True
    for t, (a, b) in enumerate(zip(res[True], res[False])):
        for i, (u, v) in enumerate(zip(a, b)):
            assert torch.equal(u, v), (t, i)
    assert abs(1 - res[True][-1][2].float().mean().item() - 0.6) < 1e-3


@pytest.mark.parametrize("running_average", [True, False])
def test_structured_callback_fused_step_equals_unfused(running_average):
    """MagnitudePruningCallback with a channel mask: the 3-launch route (partials, one parameter kernel, apply)
    equals the 9-launch route — magnitudes, masks, outputs, gradients — incl. steps without a mask refresh."""
    import importlib
    sp = importlib.import_module("qsparse_b200.sparse")
    res = {}
    for fuse in (True, False):
        sp.FUSE_PRUNE_STEP = fuse
        try:
            g = torch.Generator(device="cuda").manual_seed(5)
            cb = sp.MagnitudePruningCallback(running_average=running_average, mask_refresh_interval=2)
            cb.train()
            mask = torch.nn.Parameter(torch.ones(1, 48, 1, 1, dtype=torch.bool, device="cuda"), requires_grad=False)
            scale = torch.linspace(0.2, 2.0, 48, device="cuda").view(1, 48, 1, 1)
            rec = []
            for t in range(6):
                x = (torch.randn(6, 48, 9, 11, device="cuda", generator=g) * scale * (1 + 0.3 * t)).requires_grad_(True)
                y = cb(x, 0.5, mask)
                y.sum().backward()
                item = [y.detach().clone(), x.grad.clone(), mask.data.clone()]
                if running_average:
                    item.append(cb.magnitude.data.clone())
                rec.append(item)
            res[fuse] = rec
        finally:
            sp.FUSE_PRUNE_STEP = True
    for t, (a, b) in enumerate(zip(res[True], res[False])):
        for i, (u, v) in enumerate(zip(a, b)):
            assert torch.equal(u, v), (t, i)
    assert res[True][-1][2].sum().item() == 24


@pytest.mark.parametrize("fuse", [False, True])
def test_config1_eval_forward_in_a_cuda_graph(fuse):
    """The converted net's inference forward enqueues kernels and never waits for the device (no `.item()`, no host
    reads of device state), so the whole forward captures into ONE CUDA graph: replays on new inputs equal eager
    calls bit for bit.  (The training step is host-scheduled like the reference's — step counters, schedules,
    EMA indices are Python values — and is not capturable as a whole; its hot kernels are, see bench.py.)"""
    import qsparse_b200 as qs
    from benchmarks.configs import _c1_convert, _c1_net
    qs.set_qsparse_options(log_on_created=False)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    with contextlib.redirect_stdout(io.StringIO()):
        model = _c1_convert(torch, qs, _c1_net(torch), "ScalerQuantizer", fuse).to(dev)
        opt = torch.optim.SGD(model.parameters(), lr=0.01)
        gen = torch.Generator(device=dev).manual_seed(3)
        model.train()
        for _ in range(64):                                  # past timeout 10 and the pruning ramp 20..50
            x = torch.randn(64, 1, 28, 28, device=dev, generator=gen)
            y = torch.randint(0, 10, (64,), device=dev, generator=gen)
            opt.zero_grad()
            torch.nn.functional.nll_loss(model(x), y).backward()
            opt.step()
        model.eval()
        static_x = torch.randn(64, 1, 28, 28, device=dev, generator=gen)
        with torch.no_grad():
            eager0 = model(static_x).clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                model(static_x)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_y = model(static_x)
            graph.replay()
            assert torch.equal(static_y.view(torch.int32), eager0.view(torch.int32))
            for _ in range(3):
                xn = torch.randn(64, 1, 28, 28, device=dev, generator=gen)
                static_x.copy_(xn)
                graph.replay()
                assert torch.equal(static_y.view(torch.int32), model(xn).view(torch.int32))


def _c1_train_setup(kind, seed=11):
    import qsparse_b200 as qs
    from benchmarks.configs import _c1_convert, _c1_net
    dev = torch.device("cuda:0")
    with contextlib.redirect_stdout(io.StringIO()):
        model = _c1_convert(torch, qs, _c1_net(torch), kind, True).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.01)
    gen = torch.Generator(device=dev).manual_seed(seed)
    xs = torch.randn(70, 64, 1, 28, 28, device=dev, generator=gen)
    ys = torch.randint(0, 10, (70, 64), device=dev, generator=gen)
    return model, opt, xs, ys


@pytest.mark.parametrize("kind", ["ScalerQuantizer", "DecimalQuantizer"])
def test_config1_training_step_as_one_cuda_graph(kind):
    """qsparse_b200.GraphedTrainStep: the steady-state training step of the converted MNIST net (forward, backward,
    optimizer step; 8 quantize and 2 fused prune->quantize sites) captured into ONE CUDA graph whose step indices
    live on the device.  60 eager steps + 10 steps of which 6 are graph replays must leave the SAME parameters,
    masks, scales, magnitudes and counters — device and host side — as 70 eager steps, bit for bit."""
    import qsparse_b200 as qs
    from qsparse_b200 import graphs
    qs.set_qsparse_options(log_on_created=False)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    F = torch.nn.functional

    def run(graph_from):
        model, opt, xs, ys = _c1_train_setup(kind)
        static_x, static_y = xs[0].clone(), ys[0].clone()
        model.train()

        def step():
            opt.zero_grad(set_to_none=True)
            loss = F.nll_loss(model(static_x), static_y)
            loss.backward()
            opt.step()
            return loss

        losses = []
        with contextlib.redirect_stdout(io.StringIO()):
            i = 0
            while i < 70:
                if graph_from is not None and i == graph_from:
                    feed = iter(range(i, i + 3))

                    def warm_step():
                        j = next(feed)
                        static_x.copy_(xs[j]); static_y.copy_(ys[j])
                        return step()

                    # the warm-up steps are real steps on real batches; the captured function is `step`
                    gs = _graphed(model, step, warm_step)
                    i += 3
                    for _ in range(6):
                        static_x.copy_(xs[i]); static_y.copy_(ys[i])
                        losses.append(gs.replay().clone())
                        i += 1
                    gs.sync_host()
                    continue
                static_x.copy_(xs[i]); static_y.copy_(ys[i])
                losses.append(step().detach().clone())
                i += 1
        state = {k: v.clone() for k, v in model.state_dict().items()}
        host = [int(cb.t) for cb in model.modules() if isinstance(getattr(cb, "t", None), int)]
        fused = [m.fused_steps for m in model.modules() if hasattr(m, "fused_steps")]
        return state, host, torch.stack(losses), fused

    def _graphed(model, step, warm_step):
        """GraphedTrainStep with distinct warm-up batches: warm up through `warm_step`, capture `step`"""
        class _Two:
            def __init__(self):
                self.n = 0

            def __call__(self):
                self.n += 1
                return warm_step() if self.n <= 3 else step()
        return graphs.GraphedTrainStep(model, _Two(), warmup=3)

    ref_state, ref_host, ref_losses, ref_fused = run(None)
    got_state, got_host, got_losses, got_fused = run(60)
    assert ref_host == got_host and ref_fused == got_fused
    assert sorted(ref_state) == sorted(got_state)
    for k in ref_state:
        a, b = ref_state[k], got_state[k]
        assert torch.equal(a.view(torch.uint8) if a.dtype == torch.bool else a, b.view(torch.uint8) if b.dtype == torch.bool else b), k
    # the warm-up steps' losses are not recorded in the graphed run: compare the others
    assert torch.equal(ref_losses[:60], got_losses[:60]) and torch.equal(ref_losses[63:69], got_losses[60:66])


def test_graph_mode_refuses_what_it_cannot_capture():
    """Routes that index a running mean by a HOST step counter raise NotCapturable in graph mode instead of freezing
    the index into the graph; so does a model that is not in its steady state yet."""
    import qsparse_b200 as qs
    from qsparse_b200 import graphs
    qs.set_qsparse_options(log_on_created=False)
    dev = torch.device("cuda:0")
    with contextlib.redirect_stdout(io.StringIO()):
        # l0 importance (a data-dependent gate between two statistics) still takes a host step index
        from qsparse_b200.sparse import MagnitudePruningCallback
        lin = qs.prune(nn.Linear(64, 32), sparsity=0.5, dimensions={0, 1}, start=1, interval=1, repetition=1,
                       callback=MagnitudePruningCallback(l0=True)).to(dev)
        lin.train()
        x = torch.randn(8, 64, device=dev)
        for _ in range(4):
            lin(x)
        with graphs.graph_mode(), pytest.raises(graphs.NotCapturable):
            lin(x)
        fresh = qs.quantize(bits=8, channelwise=-1, timeout=100).to(dev)
        fresh.train()
        fresh(x)
        with pytest.raises(graphs.NotCapturable):
            graphs.GraphedTrainStep(fresh, lambda: fresh(x))


@pytest.mark.parametrize("fuse", [False, True])
def test_graphed_step_covers_the_stock_routes(fuse):
    """GraphedTrainStep over a net that exercises every capturable route at once — generic per-channel
    Decimal / Scaler estimation (`qsb_scale_ema_at`), the row-resident weight kernel (`qsb_row_quant_fused_at`,
    symmetric and Adaptive), per-channel Adaptive estimation on activations (`qsb_lines_ema_at`), the percentile
    estimator, a stand-alone structured prune layer (the one-launch step with its own `t` as the device counter), a
    quantize(prune(layer)) weight chain whose mask is frozen (plain and fused), an unstructured running-average prune
    layer (the one-pass step K9 through `qsb_prune_unstructured_step_batched_at`) — 12 eager steps + 3 warm-up + 5
    replays against 20 eager steps: every parameter, mask, scale, line, magnitude and counter bit-equal."""
    import qsparse_b200 as qs
    from qsparse_b200 import graphs
    from qsparse_b200.quantize import AdaptiveQuantizer, DecimalQuantizer, PercentileQuantizer, ScalerQuantizer
    from qsparse_b200.sparse import MagnitudePruningCallback
    qs.set_qsparse_options(log_on_created=False)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    dev = torch.device("cuda:0")
    F = torch.nn.functional

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.in_q = qs.quantize(bits=8, channelwise=-1, timeout=2, callback=PercentileQuantizer(99.0))
            self.c1 = qs.quantize(nn.Conv2d(3, 16, 3, padding=1), bits=8, channelwise=0, timeout=2,
                                  callback=DecimalQuantizer())                    # rows of 27: generic estimation
            self.c2 = qs.quantize(nn.Conv2d(16, 16, 3, padding=1), bits=8, channelwise=0, timeout=2)   # K8, Scaler
            self.c3 = qs.quantize(nn.Conv2d(16, 16, 3, padding=1), bits=8, channelwise=1, timeout=2,
                                  callback=ScalerQuantizer())                     # per input channel: generic
            self.act_q = qs.quantize(bits=8, channelwise=1, timeout=2, callback=AdaptiveQuantizer())
            self.p = qs.prune(sparsity=0.5, dimensions={1}, start=2, interval=1, repetition=2)
            self.t_q = qs.quantize(bits=8, channelwise=-1, timeout=3, callback=DecimalQuantizer())
            self.fc0 = qs.prune(nn.Linear(16 * 8 * 8, 64), sparsity=0.5, dimensions={0, 1}, start=2, interval=1,
                                repetition=2)                                     # running-average, every step: K9
            self.fc = qs.quantize(
                qs.prune(nn.Linear(64, 32), sparsity=0.5, dimensions={0, 1}, start=1, interval=1, repetition=1,
                         callback=MagnitudePruningCallback(stop_mask_refresh=4)),
                bits=4, channelwise=0, timeout=2, callback=AdaptiveQuantizer())
            self.fc2 = qs.quantize(nn.Linear(32, 10), bits=8, channelwise=0, timeout=2, callback=AdaptiveQuantizer())

        def forward(self, x):
            x = F.relu(self.c1(self.in_q(x)))
            x = self.act_q(F.relu(self.c2(x)))
            x = self.t_q(self.p(F.relu(self.c3(x))))
            return self.fc2(F.relu(self.fc(F.relu(self.fc0(x.flatten(1))))))

    def run(graph_from):
        torch.manual_seed(5)
        with contextlib.redirect_stdout(io.StringIO()):
            model = Net().to(dev)
            if fuse:
                qs.fuse_prune_quantize(model)
        opt = torch.optim.SGD(model.parameters(), lr=0.05)
        gen = torch.Generator(device=dev).manual_seed(9)
        xs = torch.randn(20, 16, 3, 8, 8, device=dev, generator=gen)
        ys = torch.randint(0, 10, (20, 16), device=dev, generator=gen)
        sx, sy = xs[0].clone(), ys[0].clone()
        model.train()

        def step():
            opt.zero_grad(set_to_none=True)
            loss = F.cross_entropy(model(sx), sy)
            loss.backward()
            opt.step()
            return loss

        with contextlib.redirect_stdout(io.StringIO()):
            i = 0
            while i < 20:
                if graph_from is not None and i == graph_from:
                    feed = iter(range(i, i + 3))

                    class _Two:
                        n = 0

                        def __call__(self):
                            self.n += 1
                            if self.n <= 3:
                                j = next(feed)
                                sx.copy_(xs[j]); sy.copy_(ys[j])
                            return step()
                    gs = graphs.GraphedTrainStep(model, _Two(), warmup=3)
                    i += 3
                    for _ in range(5):
                        sx.copy_(xs[i]); sy.copy_(ys[i])
                        gs.replay()
                        i += 1
                    gs.sync_host()
                    continue
                sx.copy_(xs[i]); sy.copy_(ys[i])
                step()
                i += 1
        state = {k: v.clone() for k, v in model.state_dict().items()}
        host = [graphs.host_index(m) for m in model.modules() if hasattr(m, "optimize") and hasattr(m, "t")]
        return state, host

    ref_state, ref_host = run(None)
    got_state, got_host = run(12)
    assert ref_host == got_host
    assert sorted(ref_state) == sorted(got_state)
    bad = [k for k in ref_state
           if ref_state[k].dtype != got_state[k].dtype or not torch.equal(
               ref_state[k].view(torch.uint8) if ref_state[k].dtype == torch.bool else ref_state[k],
               got_state[k].view(torch.uint8) if got_state[k].dtype == torch.bool else got_state[k])]
    assert not bad, bad


def test_prune_quantize_module_in_a_cuda_graph():
    """fused.PruneQuantize (the module bench.py times as `module_api`): 4 eager steps + 2 warm-up + 4 replays equal
    10 eager steps bit for bit (magnitude, mask, scale, outputs, input gradients, host counters)."""
    from qsparse_b200 import graphs
    from qsparse_b200.fused import PruneQuantize
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(4)
    xs = torch.relu(torch.randn(10, 8, 32, 14, 14, device=dev, generator=gen))
    gs_ = torch.randn(10, 8, 32, 14, 14, device=dev, generator=gen)

    def run(graph):
        layer = PruneQuantize(sparsity=0.75, bits=8).train()
        sx = xs[0].clone().requires_grad_(True)
        sg = gs_[0].clone()
        outs = []

        def step():
            sx.grad = None
            y = layer(sx)
            y.backward(sg)
            return y

        i = 0
        while i < 10:
            with torch.no_grad():
                sx.copy_(xs[i]); sg.copy_(gs_[i])
            if graph and i == 4:
                feed = iter((4, 5))

                class _Two:
                    n = 0

                    def __call__(self):
                        self.n += 1
                        if self.n <= 2:
                            j = next(feed)
                            with torch.no_grad():
                                sx.copy_(xs[j]); sg.copy_(gs_[j])
                        return step()
                g = graphs.GraphedTrainStep(layer, _Two(), warmup=2)
                i = 6
                for _ in range(4):
                    with torch.no_grad():
                        sx.copy_(xs[i]); sg.copy_(gs_[i])
                    y = g.replay()
                    outs.append((y.detach().clone(), sx.grad.clone()))
                    i += 1
                g.sync_host()
                continue
            y = step().detach()       # (a live autograd graph of an eager step would pin sx's AccumulateGrad node
            if i >= 6:                # to the default stream and invalidate the capture)
                outs.append((y.clone(), sx.grad.clone()))
            del y
            i += 1
        return layer, outs

    ref, ref_outs = run(False)
    got, got_outs = run(True)
    assert (ref.t_prune, ref.t_quant) == (got.t_prune, got.t_quant) == (10, 10)
    for name in ("magnitude", "mask", "scale"):
        a, b = getattr(ref, name).data, getattr(got, name).data
        assert torch.equal(a.view(torch.uint8) if a.dtype == torch.bool else a.view(torch.int32),
                           b.view(torch.uint8) if b.dtype == torch.bool else b.view(torch.int32)), name
    assert len(ref_outs) == len(got_outs) == 4
    for (ya, ga), (yb, gb) in zip(ref_outs, got_outs):
        assert torch.equal(ya.view(torch.int32), yb.view(torch.int32)) and torch.equal(ga.view(torch.int32), gb.view(torch.int32))


def test_graphed_prune_modules_whose_counters_started_on_the_host():
    """prune()-wrapped modules built around CUDA weights but never moved with `.cuda()` keep their callback's `t`
    Parameter on the host (as in the reference); graph mode moves it to the device before it becomes a kernel
    argument.  5 eager + 2 warm-up + 4 replays equal 11 eager steps (masks, magnitudes, counters)."""
    import qsparse_b200 as qs
    qs.set_qsparse_options(log_on_created=False)
    dev = torch.device("cuda:0")

    class _W(nn.Module):
        def __init__(self, w):
            super().__init__()
            self.weight = nn.Parameter(w.clone(), requires_grad=False)

    def run(graph, batched=False):
        gen = torch.Generator(device=dev).manual_seed(21)
        ws = [torch.randn(s, device=dev, generator=gen) * 0.02 for s in ((64, 32, 3, 3), (128, 257), (3, 100_000))]
        with contextlib.redirect_stdout(io.StringIO()):
            mods = nn.ModuleList([qs.prune(_W(w), sparsity=0.6, dimensions=set(range(w.dim())), start=0, interval=1,
                                           repetition=1) for w in ws]).train()
            pruner = qs.WeightSetPruner(mods) if batched else None

            def access():
                with torch.no_grad():
                    for m in mods:
                        m._parameters["weight"].mul_(1.001)
                    if pruner is not None:
                        pruner.step()          # the whole weight set in one batched launch sequence
                    for m in mods:
                        m.weight
            for _ in range(5):
                access()
            assert not mods[0].prune.callback.t.is_cuda
            if graph:
                gs = qs.GraphedTrainStep(mods, access, warmup=2)
                for _ in range(4):
                    gs.replay()
                gs.sync_host()
            else:
                for _ in range(6):
                    access()
        return mods

    a, b = run(False), run(True)
    c = run(True, batched=True)                # ... and with WeightSetPruner.step() inside the captured step
    for ma, mb in list(zip(a, b)) + list(zip(a, c)):
        assert torch.equal(ma.prune.mask, mb.prune.mask)
        assert torch.equal(ma.prune.callback.magnitude.view(torch.int32), mb.prune.callback.magnitude.view(torch.int32))
        assert int(ma.prune.callback.t.item()) == int(mb.prune.callback.t.item()) == 11
        assert ma.prune.callback._t() == mb.prune.callback._t() == 11
        assert int(ma.prune._n_updates.item()) == int(mb.prune._n_updates.item())
