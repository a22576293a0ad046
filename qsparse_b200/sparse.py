"""Magnitude pruning on sm_100a kernels, behind qsparse's own API.

Mirrors ``qsparse/sparse.py`` of mlzxy/qsparse v2.0.1: ``prune()``, ``PruneLayer``,
``MagnitudePruningCallback`` (+ its overridable ``initialize`` / ``receive_input`` /
``update_magnitude`` / ``prune_and_update_mask``), ``UniformPruningCallback`` and
``devise_layerwise_pruning_schedule`` with the same arguments, state_dict keys and
error types.  The tensor work runs on the kernels of ``csrc/``:

* running-average magnitude      -> K3 ``qsb_reduce_stats`` (+ ``qsb_magnitude_ema_*``)
* k-th value threshold           -> K5 ``qsb_kth_value`` (exact radix select, no sort)
* mask build / apply, fwd + bwd  -> K6 ``qsb_mask_from_threshold`` / ``qsb_mask_build_apply`` /
                                    ``qsb_mask_apply``

Step counters and the current sparsity are mirrored on the host (``HostMirror``), so a
training forward never synchronises with the device (the reference does 6-8
``.item()`` calls per step).
"""
from __future__ import annotations

import copy
from argparse import ArgumentError
from typing import Callable, Iterable

import numpy as np
import torch
import torch.nn as nn

from . import _native as N
from . import graphs
from . import ops
from .imitation import imitate
from .util import HostMirror, get_option, kth_rank, logging


# Route the stock unstructured running-average prune step through the one-pass kernel (K9).  False keeps
# update_magnitude -> k-th value -> mask build + apply (same results; the tests compare the two).
FUSE_PRUNE_STEP = True
# Warm-start the exact k-th value select of that step from the previous step's threshold (same results).
SELECT_HINTS = True


class _MaskApply(torch.autograd.Function):
    """``x * mask`` (ref qsparse/sparse.py:66,116,122,263) and its gradient ``g * mask``.

    ``precomputed`` lets the fused mask-build+apply kernel hand in the forward value."""

    @staticmethod
    def forward(ctx, x, mask, precomputed=None):
        ctx.perm = ops.mask_perm(x.shape, mask.shape)
        ctx.mask = mask  # read at backward time, like the reference's saved Parameter
        if ctx.perm is not None:            # kept axes not adjacent: transpose, same kernel, transpose back
            xs, mk = _canonical(x.detach(), mask.detach(), ctx.perm)
            _, ctx.layout = ops.mask_layout(xs.shape, mk.shape)
            if precomputed is not None:
                return precomputed
            return ops.mask_apply(xs, mk, ctx.layout).permute(ops.invert_perm(ctx.perm))
        kind, layout = ops.mask_layout(x.shape, mask.shape)
        ctx.layout = layout
        if precomputed is not None:
            return precomputed
        xs = N.as_f32_contiguous(x.detach())
        return ops.mask_apply(xs, mask.detach(), layout)

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.perm is not None:
            g, mk = _canonical(grad_output, ctx.mask.detach(), ctx.perm)
            return ops.mask_apply(g, mk, ctx.layout).permute(ops.invert_perm(ctx.perm)), None, None
        g = N.as_f32_contiguous(grad_output)
        return ops.mask_apply(g, ctx.mask.detach(), ctx.layout), None, None


def _canonical(x, like, perm):
    """(x transposed so that the kept axes of ``like`` are adjacent — a contiguous fp32 copy, the permuted VIEW of
    ``like``): the view shares storage with the mask / magnitude Parameter, whose non-singleton axes keep their
    order, so the kernels' in-place updates land in the Parameter."""
    if perm is None:
        return N.as_f32_contiguous(x), like
    return N.as_f32_contiguous(x.permute(perm)), like.permute(perm)


def apply_mask(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    N.require_cuda(x, "x")
    N.require_cuda(mask, "mask")
    return _MaskApply.apply(x, mask)


class MagnitudePruningCallback(nn.Module):
    """Magnitude-based pruning, the default callback of ``prune`` (ref qsparse/sparse.py:18-122)."""

    def __init__(self, mask_refresh_interval: int = -1, stop_mask_refresh: int = float("inf"),
                 use_gradient: bool = False, running_average: bool = True, l0: bool = False,
                 forward_hook: Callable[[torch.Tensor, str], None] = None):
        super().__init__()
        self.mask_refresh_interval = mask_refresh_interval
        self.stop_mask_refresh = stop_mask_refresh
        self.use_gradient = use_gradient
        self.t = nn.Parameter(torch.full((1,), -1), requires_grad=False)
        if use_gradient and not running_average:
            raise ArgumentError(
                None, "the combination of `use_gradient=True` and `running_average=False` is not supported")
        self.running_average = running_average
        self.prev_grad_hook = None
        self.l0 = l0
        self.forward_hook = forward_hook
        self._t_mirror = HostMirror()

    # ---- host-mirrored step counter -------------------------------------------------
    def _t(self) -> int:
        return int(self._t_mirror.get(self.t))

    def _set_t(self, value: int):
        self.t.data.fill_(value)
        self._t_mirror.wrote(self.t, value)

    @property
    def initted(self) -> bool:
        return self._t() != -1

    # ---- overridable pieces, same names as the reference ----------------------------
    def initialize(self, mask: torch.Tensor):
        if self.running_average:
            self.magnitude = nn.Parameter(torch.zeros(*mask.shape, device=mask.device, dtype=torch.float),
                                          requires_grad=False)

    def receive_input(self, x: torch.Tensor):
        if self.use_gradient:
            if self.prev_grad_hook is not None:
                self.prev_grad_hook.remove()
            if x.requires_grad:
                self.prev_grad_hook = x.register_hook(lambda grad: self.update_magnitude(grad))
            else:
                logging.error("meeting no-grad tensor")
                self.prev_grad_hook = None
        else:
            self.update_magnitude(x)

    def update_magnitude(self, x):
        """mag = (t * mag + mean|x|) / (t + 1)  (ref qsparse/sparse.py:82-89)."""
        if not self.running_average:
            return
        with torch.no_grad():
            N.require_cuda(x, "x")
            mag = self.magnitude.data
            xs, mag = _canonical(x.detach(), mag, ops.mask_perm(x.shape, mag.shape))
            t = self._t()
            kind, layout = ops.mask_layout(xs.shape, mag.shape)
            if kind == "element":
                tensor_min = None
                if self.l0:  # `x.min() == 0` gate (sparse.py:85), evaluated on the device
                    tensor_min = ops.reduce_stats(xs, (1, 1, xs.numel()), nnz=True)["tensor_min"]
                ops.magnitude_ema_full_(mag, xs, t, tensor_min, self.l0)
            else:
                stats = ops.reduce_stats(xs, layout, abssum=True, nnz=self.l0)
                ops.magnitude_ema_reduced_(mag, stats, float(layout[0] * layout[2]), t, self.l0)

    def prune_and_update_mask(self, x: torch.Tensor, sparsity: float, mask: torch.Tensor) -> torch.Tensor:
        """threshold = sorted(importance)[idx + 1]; mask = importance >= threshold; x * mask
        (ref qsparse/sparse.py:58-66, qsparse/util.py:103-117)."""
        N.require_cuda(x, "x")
        with torch.no_grad():
            perm = ops.mask_perm(x.shape, mask.shape)
            xs, mask_c = _canonical(x.detach(), mask.data, perm)
            kind, layout = ops.mask_layout(xs.shape, mask_c.shape)
            n = mask.numel()
            k = kth_rank(sparsity, n)
            if k >= n:
                raise IndexError(f"index {k} is out of bounds for dimension 0 with size {n}")
            take_abs = False
            if self.running_average:
                importance = self.magnitude.data
            elif kind == "element":
                importance, take_abs = xs, True  # importance = |x| itself, never materialised
            else:
                s = ops.reduce_stats(xs, layout, abssum=True)["abssum"]
                importance = (s / float(layout[0] * layout[2])).float()
            thr = ops.kth_value(importance, k, take_abs)
            if kind == "element":
                out = ops.mask_build_apply(importance, thr, xs, mask.data, take_abs)
            else:
                ops.mask_from_threshold(importance, thr, mask.data, take_abs)
                out = None
        return _MaskApply.apply(x, mask, out)

    def _fused_unstructured_step(self, x, sparsity, mask, t, counter=None):
        """update_magnitude + prune_and_update_mask of the stock unstructured running-average case in ONE
        streaming pass (K9, ``qsb_prune_unstructured_step_batched``: 17 B/elem instead of 29); None when
        the case is not the stock one (a subclass, gradient / l0 importance, a structured mask, ...)."""
        if not FUSE_PRUNE_STEP or type(self) is not MagnitudePruningCallback:
            return None
        if not self.running_average or self.use_gradient or self.l0 or not hasattr(self, "magnitude"):
            return None
        if not (isinstance(x, torch.Tensor) and x.is_cuda and tuple(mask.shape) == tuple(x.shape)):
            return None
        with torch.no_grad():
            xs = N.as_f32_contiguous(x.detach())
            mag = self.magnitude.data
            out = torch.empty_like(xs)
            if not ops.prune_step_supported([mag], [xs], [mask.data], [out]):
                return None
            n = mask.numel()
            k = kth_rank(sparsity, n)
            if k >= n:
                raise IndexError(f"index {k} is out of bounds for dimension 0 with size {n}")
            # warm-started pivots: the threshold of a running-average magnitude moves slowly from step to step
            # (reset when the tensor or the rank changes, e.g. at a sparsity ramp point)
            key = (mag.data_ptr(), n, k)
            if getattr(self, "_hint_key", None) != key or self._hint.device != mag.device:
                self._hint = ops.new_select_hints(1, mag.device)
                self._hint_key = key
            # graph mode (`counter`): the kernels read the step index from the callback's own `t` Parameter
            ops.prune_unstructured_step_batched_([mag], [xs], [mask.data], [out], [k], 0 if counter is not None else t,
                                                 hints=self._hint if SELECT_HINTS else None, t_dev=counter)
        return _MaskApply.apply(x, mask, out)

    def _fused_structured_step(self, x, sparsity, mask, t, refresh, counter=None):
        """update_magnitude [+ prune_and_update_mask] of the stock structured (channel-mask) case in two
        launches — the reduction, whose last-arriving CTA finalizes and does magnitude EMA, k-th value by rank
        counting and the mask; then mask apply — instead of nine (reduce + finalize, EMA, a 4-launch select over C values,
        mask build, apply).  None when the case is not the stock one."""
        if not FUSE_PRUNE_STEP or type(self) is not MagnitudePruningCallback:
            return None
        if self.use_gradient or self.l0 or (self.running_average and not hasattr(self, "magnitude")):
            return None
        if not (self.running_average or refresh):
            return None                                   # nothing to update: plain mask apply
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == mask.dim()):
            return None
        if ops.mask_perm(x.shape, mask.shape) is not None:
            return None                                   # non-adjacent kept axes: the transposing route
        kind, layout = ops.mask_layout(x.shape, mask.shape)
        outer, ch, inner = layout
        if kind != "channel" or ch != mask.numel() or ch < 2 or ch > 2048 or outer * inner < 64:
            return None
        with torch.no_grad():
            xs = N.as_f32_contiguous(x.detach())
            k = kth_rank(sparsity, ch) if refresh else 0
            if refresh and k >= ch:
                raise IndexError(f"index {k} is out of bounds for dimension 0 with size {ch}")
            if self.running_average:
                magnitude, mode = self.magnitude.data.view(-1), 1
            else:
                magnitude, mode = torch.empty(ch, dtype=torch.float32, device=x.device), 2
            dummy = torch.zeros(1, dtype=torch.float32, device=x.device)
            if counter is not None:
                # graph mode: the kernel reads the step index from `counter` (the callback's own `t` Parameter),
                # derives the refresh gate from it and advances it
                if ch > ops.FUSED_STEP_MAX_CHANNELS:
                    return None
                ops.reduce_prune_quant_step(xs, layout, magnitude, mask.data.view(-1), dummy, None,
                                            float(outer * inner), 0, mode, self.mask_refresh_interval,
                                            kth_rank(sparsity, ch), 8, 0, False, step_counter=counter)
            elif ch <= ops.FUSED_STEP_MAX_CHANNELS:
                ops.reduce_prune_quant_step(xs, layout, magnitude, mask.data.view(-1), dummy, None,
                                            float(outer * inner), t, mode, refresh, k, 8, 0, False)
            else:
                ws = ops.reduce_partials(xs, layout)
                ops.prune_quant_step_params(magnitude, mask.data.view(-1), dummy, None, ws, layout,
                                            float(outer * inner), t, mode, refresh, k, 8, 0, False)
        return apply_mask(x, mask)

    def forward(self, x: torch.Tensor, sparsity: float, mask: torch.Tensor, name=""):
        if not self.training:
            return apply_mask(x, mask)
        if not self.initted:
            self.initialize(mask)
            self._set_t(0)
            if self.mask_refresh_interval <= 0:
                self.mask_refresh_interval = 1
        t = self._t()
        refresh = (sparsity >= 0 and (t % self.mask_refresh_interval == 0 and t <= self.stop_mask_refresh)
                   and (t > 0 or not self.running_average))
        if graphs.active():
            return self._graph_mode_forward(x, sparsity, mask, t, name)
        out = None
        pre = getattr(self, "_precomputed", None)
        if pre is not None:
            # this step's magnitude / mask / output were computed for the whole weight set in one batched launch
            # sequence (WeightSetPruner.step); valid for exactly this step of this tensor
            self._precomputed = None
            if pre[0] == t and pre[1] is x and refresh and t < self.stop_mask_refresh:
                out = _MaskApply.apply(x, mask, pre[2])
        if out is None and t < self.stop_mask_refresh:
            if refresh:
                out = self._fused_unstructured_step(x, sparsity, mask, t)
            if out is None:
                out = self._fused_structured_step(x, sparsity, mask, t, refresh)
        if out is None:
            if t < self.stop_mask_refresh:
                self.receive_input(x)
            out = self.prune_and_update_mask(x, sparsity, mask) if refresh else apply_mask(x, mask)
        self.t.data.add_(1)
        self._t_mirror.wrote(self.t, t + 1)
        if self.forward_hook is not None:
            self.forward_hook(mask, name)
        return out


    def _graph_mode_forward(self, x, sparsity, mask, t, name):
        """One step whose index lives on the device (qsparse_b200.graphs): only the stock structured route, with
        the default never-stopping mask refresh."""
        pre, self._precomputed = getattr(self, "_precomputed", None), None
        if pre is not None and pre[0] == t and pre[1] is x and t > 0 and t < self.stop_mask_refresh:
            # WeightSetPruner.step() already did this step for the whole weight set (same graph, earlier node)
            out = _MaskApply.apply(x, mask, pre[2])
            graphs.callback_counter(self, x.device).add_(1)
            self._t_mirror.wrote(self.t, t + 1)
            if self.forward_hook is not None:
                self.forward_hook(mask, name)
            return out
        if t >= self.stop_mask_refresh:
            # magnitude and mask are frozen for good (t only grows): the step is a mask apply and a counter
            out = apply_mask(x, mask)
            self.t.data.add_(1)
            self._t_mirror.wrote(self.t, t + 1)
            if self.forward_hook is not None:
                self.forward_hook(mask, name)
            return out
        if self.stop_mask_refresh != float("inf") or sparsity < 0 or kth_rank(sparsity, mask.numel()) >= mask.numel():
            raise graphs.NotCapturable("a prune callback that will stop refreshing its mask later / an out-of-range "
                                       "sparsity")
        if self.mask_refresh_interval == 1 and t > 0:
            out = self._fused_unstructured_step(x, sparsity, mask, t,
                                                counter=graphs.callback_counter(self, x.device))   # full-size masks (K9)
            if out is not None:
                self.t.data.add_(1)
                self._t_mirror.wrote(self.t, t + 1)
                if self.forward_hook is not None:
                    self.forward_hook(mask, name)
                return out
        out = self._fused_structured_step(x, sparsity, mask, t, True, counter=graphs.callback_counter(self, x.device))
        if out is None:
            graphs.require_eager("this prune callback (the stock channel-mask step and the stock full-size-mask step "
                                 "with a refresh every step are capturable)")
        self._t_mirror.wrote(self.t, t + 1)                 # the kernel advanced self.t itself
        if self.forward_hook is not None:
            self.forward_hook(mask, name)
        return out


class UniformPruningCallback(MagnitudePruningCallback):
    """Unstructured uniform random pruning; never re-activates pruned positions
    (ref qsparse/sparse.py:125-152).  The random choice uses numpy's global RNG on the
    host exactly like the reference (API compatibility; not a bandwidth path).

    ``device_rng=True`` (an extension; default off) draws the positions with torch's CUDA generator
    instead (``torch.randperm`` over the surviving positions): the same distribution and the same invariants (exact
    budget, never re-activates), no O(n) host round trip — ``np.random.choice(range(n))`` takes seconds at the 64 Mi
    elements of BASELINE config 4 — but not the reference's random stream."""

    def __init__(self, *args, device_rng: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        self.device_rng = device_rng

    def initialize(self, mask: torch.Tensor):
        pass

    def receive_input(self, x: torch.Tensor):
        pass

    def prune_and_update_mask(self, x: torch.Tensor, sparsity: float, mask: torch.Tensor) -> torch.Tensor:
        cur_sparsity = (~mask).sum().item() / mask.numel()
        if cur_sparsity > sparsity:
            logging.warning("sparsity is decreasing, which shall not happen")
        budget = int(round((sparsity - cur_sparsity) * np.prod(mask.shape)))
        if self.device_rng:
            flat = mask.data.view(-1)
            alive = flat.nonzero().squeeze(1)
            flat[alive[torch.randperm(alive.numel(), device=mask.device)[:max(budget, 0)]]] = False
            return apply_mask(x, mask)
        slots = mask.nonzero(as_tuple=True)
        chosen = np.random.choice(range(len(slots[0])), size=budget, replace=False)
        mask.data[tuple(slot[chosen] for slot in slots)] = False
        return apply_mask(x, mask)


class PruneLayer(nn.Module):
    """Prune the input tensor on a schedule (ref qsparse/sparse.py:157-273).

    ``state_dict`` keys as in the reference: ``mask`` (bool, input shape with the
    non-pruned axes set to 1), ``_n_updates`` (int32 [1]), ``_cur_sparsity`` (fp32 [1]),
    ``callback.t`` (int64 [1]), ``callback.magnitude`` (fp32, mask shape)."""

    def __str__(self):
        return (f"PruneLayer(sparsity={self.sparsity}, start={self.start}, interval={self.interval}, "
                f"repetition={self.repetition}, dimensions={self.dimensions})")

    def __repr__(self):
        return str(self)

    def __init__(self, sparsity: float = 0.5, dimensions: Iterable[int] = {1},
                 callback: MagnitudePruningCallback = None, start: int = 1000, interval: int = 1000,
                 repetition: int = 4, rampup: bool = False, name=""):
        super().__init__()
        if get_option("log_on_created"):
            logging.warning(f"[Prune{name if name == '' else f' @ {name}'}] start = {start} interval = {interval} "
                            f"repetition = {repetition} sparsity = {sparsity} dimensions = {dimensions}")
        self.schedules = [start + interval * ((1 if rampup else 0) + i) for i in range(repetition)]
        self.start = start
        self.interval = interval
        self.repetition = repetition
        self.sparsity = sparsity
        self.name = name
        self.callback = callback if callback is not None else MagnitudePruningCallback()
        self.rampup_interval = 0 if rampup else interval
        self.dimensions = set(dimensions)
        for key in ("mask", "_n_updates", "_cur_sparsity"):  # shape-less placeholders until the first forward
            self.register_parameter(key, nn.Parameter(torch.tensor(-1, dtype=torch.int), requires_grad=False))
        self._n_mirror = HostMirror()
        self._s_mirror = HostMirror()

    @property
    def initted(self) -> bool:
        return self._n_mirror.get(self._n_updates) != -1

    def _allocate(self, x: torch.Tensor):
        assert len(x.shape) > 1
        N.require_cuda(x, "x")
        mask_shape = [s if i in self.dimensions else 1 for i, s in enumerate(x.shape)]
        self.mask = nn.Parameter(torch.ones(*mask_shape, dtype=torch.bool, device=x.device), requires_grad=False)
        if self.mask.numel() == 1:
            logging.warn(f"the mask shape of {self.name} is {tuple(self.mask.shape)}, which is not prunable")
        self._n_updates = nn.Parameter(torch.zeros(1, dtype=torch.int, device=x.device), requires_grad=False)
        self._cur_sparsity = nn.Parameter(torch.zeros(1, device=x.device), requires_grad=False)
        self._n_mirror.wrote(self._n_updates, 0)
        self._s_mirror.wrote(self._cur_sparsity, 0.0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.initted:
            self._allocate(x)
        n = self._n_mirror.get(self._n_updates)
        if (n in self.schedules) and self.training:
            # cubic ramp, stored in an fp32 parameter (ref qsparse/sparse.py:252-257)
            ratio = (1.0 - (n - self.start + self.rampup_interval) / (self.interval * self.repetition)) ** 3
            value = self.sparsity * (1 - ratio)
            self._cur_sparsity[0] = value
            self._s_mirror.wrote(self._cur_sparsity, float(np.float32(value)))
            logging.warning(f"[Prune{self.name if self.name == '' else f' @ {self.name}'}] [Step {n}] "
                            f"pruned {float(np.float32(value)):.02f}")
        if not self.training or self.mask.numel() == 1:
            return apply_mask(x, self.mask)
        if n >= self.start:
            if n == self.start:
                logging.warning(f"Start pruning at {self.name} @ {n}")
            out = self.callback(x, self._s_mirror.get(self._cur_sparsity), mask=self.mask, name=self.name)
        else:
            out = x
        self._n_updates.data.add_(1)
        self._n_mirror.wrote(self._n_updates, n + 1)
        return out


class WeightSetPruner:
    """EXTENSION (SURVEY 8e "weights"): the unstructured running-average prune step of EVERY pruned
    weight of a model in one batched launch sequence instead of one sequence per layer.

    Every ``prune(layer, dimensions=all axes)`` module runs, on each access of its ``weight``, magnitude EMA +
    exact k-th value + mask + apply for its own tensor (``MagnitudePruningCallback.forward``): 7 launches per
    layer and step, 1.9 ms for the 29 layers of BASELINE config 4 although the work is 0.2 ms.  The weights
    only change at ``optimizer.step()``, so the whole set can be done up front::

        pruner = WeightSetPruner(model)
        for batch in data:
            pruner.step()            # one streaming pass over all pruned weights (K9, warm-started pivots)
            loss = model(batch) ...  # each layer's PruneLayer picks its precomputed mask / output up

    Results are bit-identical to the per-layer path: the same kernels run with the same arguments; a layer whose
    state does not allow it this step (not started, ramp point, eval, custom callback, first step) is simply left
    to its own forward.  ``step()`` must run after the weights were last modified and before the forward."""

    def __init__(self, model: nn.Module):
        self.layers = []
        for mod in model.modules():
            p = getattr(mod, "prune", None)
            if isinstance(p, PruneLayer) and "weight" in getattr(mod, "_parameters", {}):
                self.layers.append((mod, p))
        self._hints = None
        self._hint_key = None

    def _eligible(self, mod, p):
        cb = p.callback
        w = mod._parameters["weight"]
        if not (p.training and cb.training and type(cb) is MagnitudePruningCallback and FUSE_PRUNE_STEP):
            return None
        if not p.initted or not cb.initted or not cb.running_average or cb.use_gradient or cb.l0:
            return None
        if not hasattr(cb, "magnitude") or tuple(p.mask.shape) != tuple(w.shape):
            return None
        n = p._n_mirror.get(p._n_updates)
        if n < p.start or n in p.schedules:
            return None
        t = cb._t()
        sparsity = p._s_mirror.get(p._cur_sparsity)
        refresh = (sparsity >= 0 and (t % cb.mask_refresh_interval == 0 and t <= cb.stop_mask_refresh) and t > 0)
        if not refresh or t >= cb.stop_mask_refresh:
            return None
        if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()):
            return None
        k = kth_rank(sparsity, w.numel())
        if k >= w.numel():
            return None
        return w, cb, t, k

    @torch.no_grad()
    def step(self) -> int:
        """Precompute this step for every eligible layer; returns how many layers were batched.
        In graph mode (``qsparse_b200.graphs``) the one batched launch sequence reads the step index from the first
        layer's callback counter, so all batched layers must be at the same step."""
        todo = []
        for mod, p in self.layers:
            e = self._eligible(mod, p)
            if e is not None:
                todo.append((p,) + e)
        if not todo:
            return 0
        # all layers of one batch share the callback step index only if they started together: group by t
        done = 0
        by_t = {}
        for item in todo:
            by_t.setdefault(item[3], []).append(item)
        if graphs.active() and len(by_t) != 1:
            raise graphs.NotCapturable("WeightSetPruner over layers at different callback steps")
        for t, items in by_t.items():
            ws = [it[1].detach() for it in items]
            mags = [it[2].magnitude.data for it in items]
            masks = [it[0].mask.data for it in items]
            outs = [torch.empty_like(w) for w in ws]
            ks = [it[4] for it in items]
            if not ops.prune_step_supported(mags, ws, masks, outs):
                continue
            key = tuple((m.data_ptr(), k) for m, k in zip(mags, ks))
            if self._hint_key != key:
                self._hints = ops.new_select_hints(len(items), ws[0].device)
                self._hint_key = key
            counter = graphs.callback_counter(items[0][2], ws[0].device) if graphs.active() else None
            ops.prune_unstructured_step_batched_(mags, ws, masks, outs, ks, 0 if counter is not None else t,
                                                 hints=self._hints if SELECT_HINTS and len(by_t) == 1 else None,
                                                 t_dev=counter)
            for it, o in zip(items, outs):
                it[2]._precomputed = (t, it[1], o)
            done += len(items)
        return done


def prune(inp: nn.Module = None, sparsity: float = 0.5, dimensions: Iterable[int] = {1},
          callback: MagnitudePruningCallback = None, start: int = 1000, interval: int = 1000,
          repetition: int = 4, rampup: bool = False, name="") -> nn.Module:
    """Create a ``PruneLayer`` (no input module) or wrap the weight of ``inp`` with one
    (ref qsparse/sparse.py:276-339)."""
    callback = callback or MagnitudePruningCallback()
    kwargs = dict(start=int(start), sparsity=sparsity, interval=int(interval), repetition=repetition,
                  rampup=rampup, name=name, callback=callback, dimensions=dimensions)
    if inp is None:
        layer = PruneLayer(**kwargs)
        setattr(layer, "_kwargs", kwargs)
        return layer
    if isinstance(inp, nn.Module):
        return imitate(inp, "prune", PruneLayer(**kwargs))
    raise ValueError(f"{inp} is not a valid argument for prune")


def devise_layerwise_pruning_schedule(net: nn.Module, start: int = 1, interval: int = 10,
                                      mask_refresh_interval: int = 1, inplace=False, fix_ramp: bool = False):
    """Stagger the start of every PruneLayer, attribute for attribute as the reference
    does (ref qsparse/sparse.py:343-359).  Note that the reference leaves
    ``rampup_interval`` untouched, which makes the ramp formula of ``PruneLayer.forward``
    (sparse.py:252-257) evaluate ``(1 - old_interval / new_interval)^3`` at the layer's single
    schedule point — a target sparsity far outside [0, 1] whenever the layer was created with a
    different interval.  Default: that behaviour is kept, not fixed.

    ``fix_ramp=True`` (extension, opt-in): also set ``rampup_interval = interval`` on every layer, so
    the one-shot schedule reaches exactly the layer's configured ``sparsity``."""
    if not inplace:
        net = copy.deepcopy(net)
    layers = [m for m in net.modules() if isinstance(m, PruneLayer)]
    weight_only = all(p.name.endswith(".prune") for p in layers)
    for layer in layers:
        layer.start = start
        layer.interval = interval
        layer.repetition = 1
        layer.schedules = [start]
        layer.callback.mask_refresh_interval = mask_refresh_interval
        layer.callback.stop_mask_refresh = interval
        if fix_ramp:
            layer.rampup_interval = interval
        if weight_only:
            layer.callback.running_average = False
        start += interval + 1
    logging.danger(f"Pruning stops at iteration - {start}")
    return net
