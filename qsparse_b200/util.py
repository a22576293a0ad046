"""Host utilities of the quantize/prune path (ref qsparse/util.py).

Options, naming, checkpoint pre-loading and the two tensor helpers the prune path
is built on.  The tensor helpers run on the CUDA kernels (K3 reductions, K5
select); everything else is host-only bookkeeping.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _native as N
from . import ops

_options_ = {"log_on_created": True, "log_during_train": True}


def set_options(log_on_created: Optional[bool] = None, log_during_train: Optional[bool] = None):
    """Update the two logging switches; ``None`` leaves a switch alone (ref qsparse/util.py:13-26)."""
    for key, val in (("log_on_created", log_on_created), ("log_during_train", log_during_train)):
        if val is not None:
            _options_[key] = val


def get_option(key: str):
    """ref qsparse/util.py:29-40"""
    assert key in ("log_on_created", "log_during_train")
    return _options_[key]


def auto_name_prune_quantize_layers(net: nn.Module) -> nn.Module:
    """Name every Prune/Quantize layer after its module path (ref qsparse/util.py:43-60)."""
    from .quantize import QuantizeLayer
    from .sparse import PruneLayer

    for path, mod in net.named_modules():
        if isinstance(mod, (PruneLayer, QuantizeLayer)):
            mod.name = path
    return net


def nn_module(mod: nn.Module) -> nn.Module:
    """Unwrap ``nn.DataParallel``-style containers (ref qsparse/util.py:64-76)."""
    return mod.module if hasattr(mod, "module") else mod


def kth_rank(sparsity: float, n: int) -> int:
    """0-based ascending rank of the prune threshold: ``idx + 1`` with
    ``idx = max(int(sparsity * n - 1), 0)`` (ref qsparse/util.py:114-116; evaluated in
    float64 and truncated)."""
    return max(int(sparsity * n - 1), 0) + 1


def squeeze_tensor_to_shape(x: torch.Tensor, shape: List[int]) -> torch.Tensor:
    """Mean-reduce ``x`` over every axis where ``shape`` is 1 (ref qsparse/util.py:79-99).

    One K3 reduction pass (fp64 accumulation of the kept-axis sums, a single
    division) instead of the reference's chain of fp32 ``mean`` calls; agrees with
    it to a few ulp."""
    assert len(x.shape) == len(shape), "mismatch between the input tensor and mask"
    shape = [int(s) for s in shape]
    if all(int(sx) == sm for sx, sm in zip(x.shape, shape)):
        return x
    for sx, sm in zip(x.shape, shape):
        if sx != sm and sm != 1:
            raise ValueError("mismatch between the input tensor and mask")
    N.require_cuda(x, "x")
    # the reduction kernel sums |x|; callers pass x.abs() (non-negative), for a
    # signed x the sum of x itself is needed -> split by sign.
    xs = N.as_f32_contiguous(x.detach())
    perm = ops.mask_perm(xs.shape, shape)
    if perm is not None:      # kept axes not adjacent: transpose first; the [ch] result is in kept-axis order either way
        xs = xs.permute(perm).contiguous()
        _, layout = ops.mask_layout(xs.shape, [shape[p] for p in perm])
    else:
        _, layout = ops.mask_layout(xs.shape, shape)
    outer, ch, inner = layout
    count = float(outer * inner)
    pos = ops.reduce_stats(torch.clamp_min(xs, 0) if _maybe_negative(xs) else xs, layout, abssum=True)["abssum"]
    if _maybe_negative(xs):
        neg = ops.reduce_stats(torch.clamp_max(xs, 0), layout, abssum=True)["abssum"]
        pos = pos - neg
    return (pos / count).float().view(shape)


def _maybe_negative(x: torch.Tensor) -> bool:
    # Callers on the hot path use mean_abs_to_shape() below instead (no sign question,
    # no sync); this generic entry point answers it with one tiny device reduction.
    return bool((x < 0).any().item())


def mean_abs_to_shape(x: torch.Tensor, shape) -> torch.Tensor:
    """``squeeze_tensor_to_shape(x.abs(), shape)`` without materialising ``x.abs()`` and
    without a host sync: the hot-path form (ref qsparse/sparse.py:64,87)."""
    N.require_cuda(x, "x")
    xs = N.as_f32_contiguous(x.detach())
    shape = [int(s) for s in shape]
    perm = ops.mask_perm(xs.shape, shape)
    if perm is not None:
        xs = xs.permute(perm).contiguous()
        kind, layout = ops.mask_layout(xs.shape, [shape[p] for p in perm])
    else:
        kind, layout = ops.mask_layout(xs.shape, shape)
    if kind == "element" and xs.numel() == _prod(shape):
        return xs.abs().view(shape)
    outer, ch, inner = layout
    s = ops.reduce_stats(xs, layout, abssum=True)["abssum"]
    return (s / float(outer * inner)).float().view(shape)


def _prod(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


def calculate_mask_given_importance(importance: torch.Tensor, sparsity: float) -> torch.Tensor:
    """Binary mask keeping ``importance >= sorted(importance)[idx + 1]`` (ref
    qsparse/util.py:103-117).  The threshold comes from the exact radix select (K5)
    instead of a full sort; it never leaves the device."""
    N.require_cuda(importance, "importance")
    imp = N.as_f32_contiguous(importance.detach())
    n = imp.numel()
    k = kth_rank(sparsity, n)
    if k >= n:
        raise IndexError(f"index {k} is out of bounds for dimension 0 with size {n}")
    thr = ops.kth_value(imp, k)
    mask = torch.empty(imp.shape, dtype=torch.bool, device=imp.device)
    ops.mask_from_threshold(imp, thr, mask)
    return mask


def preload_qsparse_state_dict(model: nn.Module, state_dict: Dict[str, torch.Tensor]) -> nn.Module:
    """Install correctly shaped Parameters for every Prune/Quantize layer (and their
    callbacks) before ``load_state_dict`` — their shapes are only known after the first
    forward (ref qsparse/util.py:120-145)."""
    from .quantize import QuantizeLayer
    from .sparse import PruneLayer

    device = next(iter(model.parameters())).device
    names = list(state_dict.keys())
    for layer_path, layer in model.named_modules():
        if not isinstance(layer, (PruneLayer, QuantizeLayer)):
            continue
        for sub_path, sub in layer.named_modules():
            prefix = "".join(p + "." for p in (layer_path, sub_path) if p)
            for key in names:
                leaf = key[len(prefix):]
                if key.startswith(prefix) and "." not in leaf:
                    sub._parameters[leaf] = nn.Parameter(state_dict[key].to(device), requires_grad=False)
    return model


class HostMirror:
    """Host copy of a tiny device Parameter (step counter, current sparsity).

    The reference reads these back with ``.item()`` on every forward (6-8 host syncs
    per layer-step).  Here the value is tracked on the host; the device
    copy is still updated (it is what ``state_dict`` saves).  If somebody else wrote
    the Parameter — ``load_state_dict``, ``preload_qsparse_state_dict``, user code —
    its identity or version counter changes and the mirror re-reads it once."""

    __slots__ = ("_id", "_version", "_value")

    def __init__(self):
        self._id = None
        self._version = None
        self._value = None

    def get(self, param: torch.Tensor):
        if self._id != id(param) or self._version != param._version:
            self._value = param.item()  # one sync, only after an external write
            self._id, self._version = id(param), param._version
        return self._value

    def wrote(self, param: torch.Tensor, value):
        """Call after updating ``param`` in place to ``value`` ourselves."""
        self._id, self._version, self._value = id(param), param._version, value


class style:
    RED, GREEN, YELLOW, RESET = "\033[31m", "\033[32m", "\033[33m", "\033[0m"


def _printer(color: str = ""):
    def emit(msg: str):
        print(f"{color}{msg}{style.RESET}" if color else msg)

    return emit


class logging:
    """print-based logger with the reference's method names (ref qsparse/util.py:162-182)."""

    info = staticmethod(_printer())
    warn = staticmethod(_printer(style.YELLOW))
    warning = staticmethod(_printer(style.YELLOW))
    error = staticmethod(_printer(style.RED))
    danger = staticmethod(_printer(style.RED))
    exception = staticmethod(_printer(style.RED))
    debug = staticmethod(_printer(style.GREEN))
