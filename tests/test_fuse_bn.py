"""fuse_bn scenarios (the ones qsparse's tests/test_fuse.py:7-130 pins): the folded network computes the
same function in eval mode, and the batch-norms are gone.  Offline weight algebra in plain torch (no kernels of
ours), so it runs wherever the model lives; here on the CPU."""
import torch
import torch.nn as nn


def _randomise_bn(net):
    g = torch.Generator().manual_seed(0)
    for m in net.modules():
        if type(m).__name__.startswith("BatchNorm"):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g))
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.num_features, generator=g))


def _same_function(net, x, **kw):
    from qsparse_b200.fuse import fuse_bn
    _randomise_bn(net)
    net = net.eval()
    want = net(x)
    fused = fuse_bn(net, log=False, inplace=False, **kw).eval()
    got = fused(x)
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-4)
    return fused


def test_conv_linear_deconv():
    f = _same_function(nn.Sequential(nn.Conv2d(3, 8, 3, bias=False), nn.BatchNorm2d(8)), torch.randn(2, 3, 9, 9))
    assert isinstance(f, nn.Conv2d) and f.bias is not None
    f = _same_function(nn.Sequential(nn.Linear(12, 7), nn.BatchNorm1d(7)), torch.randn(5, 12))
    assert isinstance(f, nn.Linear)
    f = _same_function(nn.Sequential(nn.ConvTranspose2d(4, 6, 3), nn.BatchNorm2d(6), nn.ReLU()),
                       torch.randn(2, 4, 5, 5))
    assert "batchnorm" not in str(f).lower()


def test_nested_sequentials_and_selection():
    net = nn.Sequential(
        nn.Conv2d(3, 8, 3), nn.Sequential(nn.BatchNorm2d(8), nn.ReLU()),           # BN opens a nested container
        nn.Sequential(nn.Conv2d(8, 8, 3), nn.BatchNorm2d(8)), nn.ReLU(),
        nn.Sequential(nn.Sequential(nn.Conv2d(8, 4, 1)), nn.BatchNorm2d(4)),
        nn.Flatten(), nn.Linear(4 * 5 * 5, 6), nn.BatchNorm1d(6))
    f = _same_function(net, torch.randn(3, 3, 9, 9))
    assert "batchnorm" not in str(f).lower()
    # only the requested layer types are folded
    net = nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4), nn.Flatten(), nn.Linear(4 * 7 * 7, 5), nn.BatchNorm1d(5))
    f = _same_function(net, torch.randn(2, 3, 9, 9), layers=["Linear"])
    assert str(f).lower().count("batchnorm") == 1


def test_non_sequential_root_and_wrapper():
    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.features = nn.Sequential(nn.Conv2d(3, 6, 3), nn.BatchNorm2d(6), nn.ReLU())
            self.head = nn.Sequential(nn.Flatten(), nn.Linear(6 * 7 * 7, 4), nn.BatchNorm1d(4))

        def forward(self, x):
            return self.head(self.features(x))
    f = _same_function(Net(), torch.randn(2, 3, 9, 9))
    assert "batchnorm" not in str(f).lower()
    from qsparse_b200.fuse import fuse_bn
    wrapped = nn.DataParallel(nn.Sequential(nn.Conv2d(3, 6, 3), nn.BatchNorm2d(6)))
    fused = fuse_bn(wrapped, log=False)
    assert "batchnorm" not in str(fused).lower()
