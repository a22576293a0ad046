"""oracle/gen_golden_config1.py — TEST INFRASTRUCTURE.

BASELINE config 1: the `Net` of the reference's examples/mnist.py:17-44, converted with
8-bit quantize + 50 % channel prune (integer step-wise schedule), trained for
60 steps on synthetic MNIST-shaped batches with the UNMODIFIED reference on CPU.  Records
per-step loss, the final masks, scales and a few per-step snapshots to
tests/golden/config1_mnist.npz.  Run in the build container only (needs /root/reference).
"""
import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
STEPS, BATCH = 60, 64


class Net(nn.Module):
    """same topology as examples/mnist.py:17-44 (re-typed, not imported: the example imports torchvision)"""

    def __init__(self):
        super().__init__()
        self.conv_part = nn.Sequential(
            nn.Conv2d(1, 32, 3, 1), nn.BatchNorm2d(32), nn.ReLU(),
            nn.Conv2d(32, 64, 3, 1), nn.BatchNorm2d(64), nn.ReLU(),
            nn.MaxPool2d(2), nn.Dropout(0.25))
        self.linear_part = nn.Sequential(
            nn.Flatten(), nn.Linear(9216, 128), nn.BatchNorm1d(128), nn.ReLU(), nn.Dropout(0.5), nn.Linear(128, 10))

    def forward(self, x):
        return F.log_softmax(self.linear_part(self.conv_part(x)), dim=1)


def build(qs, callback_factory):
    """mirrors examples/mnist.py:193-199 with the integer step-wise schedule of SURVEY §8(d) config 1"""
    torch.manual_seed(1)
    model = Net()
    with contextlib.redirect_stdout(io.StringIO()):
        model = qs.convert(model, qs.prune(sparsity=0.5, dimensions={1}, start=20, interval=10, repetition=4),
                           activation_layers=[nn.ReLU], excluded_activation_layer_indexes=[(nn.ReLU, [-1])])
        model = qs.convert(model, qs.quantize(bits=8, channelwise=-1, timeout=10, callback=callback_factory()),
                           activation_layers=[nn.ReLU], weight_layers=[nn.Conv2d, nn.Linear], input=True)
    for m in model.modules():           # dropout off: its RNG differs between devices
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    return model


def data(step):
    g = torch.Generator().manual_seed(1000 + step)
    x = torch.randn(BATCH, 1, 28, 28, generator=g)
    y = torch.randint(0, 10, (BATCH,), generator=g)
    return x, y


def run(qs, callback_factory, device="cpu", record=None):
    from_layers = lambda model, cls: [(n, m) for n, m in model.named_modules() if m.__class__.__name__ == cls]
    model = build(qs, callback_factory).to(device)
    model.train()
    opt = torch.optim.Adadelta(model.parameters(), lr=1.0)
    losses, sparsities = [], []
    trace = {"mag0": [], "mask0": [], "mag1": [], "mask1": [], "qw": []}
    with contextlib.redirect_stdout(io.StringIO()):
        for step in range(STEPS):
            x, y = data(step)
            x, y = x.to(device), y.to(device)
            opt.zero_grad()
            loss = F.nll_loss(model(x), y)
            loss.backward()
            opt.step()
            losses.append(loss.item())
            players = from_layers(model, "PruneLayer")
            sparsities.append([float((~m.mask).float().mean().item()) if m.mask.dim() > 0 else 0.0
                               for _, m in players])
            for i, (_, m) in enumerate(players):
                has = hasattr(m.callback, "magnitude")
                trace[f"mag{i}"].append(m.callback.magnitude.detach().cpu().numpy().reshape(-1).copy() if has
                                        else np.zeros(m.mask.numel() if m.mask.dim() else 1, np.float32))
                trace[f"mask{i}"].append(m.mask.detach().cpu().numpy().reshape(-1).copy() if m.mask.dim()
                                         else np.ones(1, bool))
            trace["qw"].append(np.array([m.weight.detach().cpu().numpy().reshape(-1)[0] if hasattr(m, "weight")
                                         else 0.0 for _, m in from_layers(model, "QuantizeLayer")], np.float32))
    out = {"loss": np.array(losses, np.float64), "sparsity": np.array(sparsities, np.float64)}
    for k, v in trace.items():
        n = max(len(a) for a in v)
        out["trace_" + k] = np.stack([np.resize(a, n) if len(a) != n else a for a in v])
    for i, (n, m) in enumerate(from_layers(model, "PruneLayer")):
        out[f"prune{i}_mask"] = m.mask.detach().cpu().numpy()
        out[f"prune{i}_magnitude"] = m.callback.magnitude.detach().cpu().numpy()
        out[f"prune{i}_name"] = np.array(n)
    for i, (n, m) in enumerate(from_layers(model, "QuantizeLayer")):
        out[f"quant{i}_weight"] = m.weight.detach().cpu().numpy()
        out[f"quant{i}_n_updates"] = m._n_updates.detach().cpu().numpy()
        out[f"quant{i}_name"] = np.array(n)
    return out, model


def main():
    sys.path.insert(0, "/root/reference")
    import qsparse as qs
    assert qs.__file__.startswith("/root/reference")
    qs.set_qsparse_options(log_on_created=False)
    from qsparse.quantize import DecimalQuantizer, ScalerQuantizer
    g = {}
    for tag, factory in (("scaler", ScalerQuantizer), ("decimal", DecimalQuantizer)):
        out, _ = run(qs, factory)
        for k, v in out.items():
            g[f"{tag}/{k}"] = v
        print(tag, "final loss", out["loss"][-1], "sparsity", out["sparsity"][-1])
    np.savez_compressed(OUT / "config1_mnist.npz", **g)
    print("wrote", OUT / "config1_mnist.npz", sum(v.nbytes for v in g.values()) / 1e3, "KB")


if __name__ == "__main__":
    main()
