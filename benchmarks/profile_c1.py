"""cProfile of the config-1 training step through the module API (development tool): where the host time of
the launch-bound regime goes."""
import cProfile
import contextlib
import io
import pstats
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import qsparse_b200 as q  # noqa: E402
from benchmarks import configs  # noqa: E402

dev = torch.device("cuda:0")
q.set_qsparse_options(log_on_created=False)
F = torch.nn.functional
with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
    model = configs._c1_convert(torch, q, configs._c1_net(torch), "ScalerQuantizer", True).to(dev).train()
opt = torch.optim.Adadelta([p for p in model.parameters() if p.requires_grad], lr=1.0)
x = torch.randn(64, 1, 28, 28, device=dev)
y = torch.randint(0, 10, (64,), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = F.nll_loss(model(x), y)
    loss.backward()
    opt.step()


with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
    for _ in range(80):
        step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
torch.cuda.synchronize()
pr.disable()
for key in ("tottime", "cumulative"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue())
