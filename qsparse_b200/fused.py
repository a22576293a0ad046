"""Fused structured prune -> pow2 quantize (K7): the training step of
``Sequential(PruneLayer(dimensions={channel}), QuantizeLayer(channelwise=-1,
callback=DecimalQuantizer()))`` (the adjacency ``convert()`` creates,
ref qsparse/convert.py:214-217) in three launches and 20 B/elem:

    forward   reduce_prune_quant_step(x) 4 B/elem   per-channel sum|x| and max|x| in one read; the
                                                    kernel's last-arriving CTA finalizes, exchanges the
                                                    statistics row with the peer GPUs (NVLink packets)
                                                    and derives magnitude EMA, k-th threshold, mask,
                                                    abs-max of kept channels, scale EMA, decimal
              fq_pow2_fwd(x, mask)       8 B/elem   y = Q(x * mask); pruned channels are not read
    backward  ste_bwd(g, mask)           8 B/elem   gx = clamp(g) * mask

against the reference's ~60 B/elem forward.  No host synchronisation:
step counters are host integers, every derived scalar stays on the device.  With a
process group the statistics row is exchanged over peer memory inside that kernel
(parallel.P2PExchange; NCCL all-gather + a parameter kernel when peer mapping is
unavailable) and combined in rank order, so every rank derives identical parameters.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _native as N
from . import graphs
from . import ops
from .parallel import StatExchange, make_exchange
from .util import kth_rank


class _PruneQuantizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, layer):
        ctx.layer = layer
        ctx.layout = layer._layout(x)
        return layer._forward_impl(x, ctx.layout)

    @staticmethod
    def backward(ctx, grad_output):
        return ctx.layer.backward_kernel(grad_output, ctx.layout), None


class PruneQuantize(nn.Module):
    """State mirrors the two reference layers: ``magnitude`` / ``mask`` (callback.magnitude,
    PruneLayer.mask, shape [C]) and ``scale`` (QuantizeLayer.weight, shape [1, 1])."""

    def __init__(self, sparsity: float = 0.5, bits: int = 8, channel_index: int = 1, running_average: bool = True,
                 mask_refresh_interval: int = 1, group=None, mutate_grad_output: bool = False,
                 exchange_timeout_ms: int = 30000, check_every: int = 16):
        super().__init__()
        self.sparsity = sparsity
        self.bits = bits
        self.channel_index = channel_index
        self.running_average = running_average
        self.mask_refresh_interval = max(int(mask_refresh_interval), 1)
        self.group = group
        self.mutate_grad_output = mutate_grad_output  # also clamp grad_output in place (quantize.py:72)
        self.exchange_timeout_ms = exchange_timeout_ms
        self.check_every = max(int(check_every), 1)   # steps between polls of the exchange's error flag
        self.t_prune = 0   # MagnitudePruningCallback.t
        self.t_quant = 0   # DecimalQuantizer.t
        self._exchange: Optional[StatExchange] = None

    def _layout(self, x):
        return N.channel_layout(x.shape, self.channel_index)

    def _step_counter(self, device):
        """device twin of ``t_prune`` for graph mode (created outside capture by ``GraphedTrainStep``)"""
        c = getattr(self, "_t_dev", None)
        if c is None or c.device != device:
            if torch.cuda.is_current_stream_capturing():
                raise graphs.NotCapturable("PruneQuantize met its first graph-mode step during capture")
            c = torch.full((1,), int(self.t_prune), dtype=torch.int64, device=device)
            self._t_dev = c
        return c

    def _allocate(self, x, channels):
        dev = x.device
        self.magnitude = nn.Parameter(torch.zeros(channels, device=dev), requires_grad=False)
        self.mask = nn.Parameter(torch.ones(channels, dtype=torch.bool, device=dev), requires_grad=False)
        self.scale = nn.Parameter(torch.zeros(1, 1, device=dev), requires_grad=False)
        self.decimal = torch.zeros(1, device=dev)
        import torch.distributed as dist
        from .parallel import assert_equal_shards
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if multi:
            assert_equal_shards(x.numel() // channels, self.group)
        self._p2p = make_exchange(channels, dev, self.group, timeout_ms=self.exchange_timeout_ms) \
            if channels <= 2048 else None
        # NCCL all-gather path only when peer memory is unavailable (or > 2048 channels)
        self._exchange = StatExchange(channels, dev, self.group) if (multi and self._p2p is None) or channels > 2048 \
            else True

    def _forward_impl(self, x, layout):
        outer, ch, inner = layout
        xs = N.as_f32_contiguous(x.detach())
        if self._exchange is None:
            self._allocate(xs, ch)
        if self.training:
            t = self.t_prune
            refresh = (t % self.mask_refresh_interval == 0) and (t > 0 or not self.running_average)
            mode = 1 if self.running_average else 2
            k = kth_rank(self.sparsity, ch)
            if isinstance(self._exchange, StatExchange):
                # NCCL fallback: finalize into the row, all-gather, combine rows in the kernel
                graphs.require_eager("PruneQuantize over the NCCL all-gather fallback")
                ex = self._exchange
                ops.reduce_stats(xs, layout, abssum=True, absmax=True,
                                 out={"abssum": ex.row.abssum, "absmax": ex.row.absmax})
                rows, n_rows, stride = ex.gather()
                abssum0, absmax0 = ex.views(rows)
                ops.prune_quant_params(self.magnitude.data, self.mask.data, self.scale.data, self.decimal,
                                       {"abssum": abssum0, "absmax": absmax0}, float(outer * inner * n_rows), t,
                                       mode, refresh, k, self.bits, self.t_quant, True, n_rows=n_rows,
                                       row_stride_bytes=stride)
            else:
                world = self._p2p.world if self._p2p is not None else 1
                grp = self._p2p.handle if self._p2p is not None else None
                stamp = self._p2p.next_stamp() if self._p2p is not None else 1
                if graphs.active():
                    # graph mode: the step index (and the exchange stamp) come from a device counter that the
                    # kernel reads and advances itself
                    if ch > ops.FUSED_STEP_MAX_CHANNELS:
                        raise graphs.NotCapturable(f"PruneQuantize with more than {ops.FUSED_STEP_MAX_CHANNELS} channels")
                    ops.reduce_prune_quant_step(xs, layout, self.magnitude.data, self.mask.data, self.scale.data,
                                                self.decimal, float(outer * inner * world), 0, mode,
                                                self.mask_refresh_interval, k, self.bits, self.t_quant - t, True,
                                                group=grp, step_counter=self._step_counter(xs.device))
                elif ch <= ops.FUSED_STEP_MAX_CHANNELS:
                    # ONE kernel: reduction whose last CTA finalizes, exchanges and derives the parameters
                    ops.reduce_prune_quant_step(xs, layout, self.magnitude.data, self.mask.data, self.scale.data,
                                                self.decimal, float(outer * inner * world), t, mode, refresh, k,
                                                self.bits, self.t_quant, True, group=grp, step_stamp=stamp)
                else:
                    ws = ops.reduce_partials(xs, layout)
                    ops.prune_quant_step_params(self.magnitude.data, self.mask.data, self.scale.data, self.decimal,
                                                ws, layout, float(outer * inner * world), t, mode, refresh, k,
                                                self.bits, self.t_quant, True, group=grp, step_stamp=stamp)
                if self._p2p is not None and t % self.check_every == 0 and not graphs.active():
                    self._p2p.check()     # raises when an earlier step timed out waiting for a peer
            self.t_prune += 1
            self.t_quant += 1
        return ops.fq_pow2_fwd(xs, self.decimal, layout, mask=self.mask.data)

    def backward_kernel(self, g, layout):
        g = N.as_f32_contiguous(g)
        _, gx = ops.ste_bwd(g, self.decimal, True, self.bits, 0, layout, mask=self.mask.data,
                            clamp_in_place=self.mutate_grad_output, want_gx=True)
        return gx

    def forward(self, x):
        N.require_cuda(x, "x")
        return _PruneQuantizeFn.apply(x, self)


# ----------------------------------------------------------------------------- fusion pass
class _FusedLayersFn(torch.autograd.Function):
    """forward value and backward of ``QuantizeLayer(PruneLayer(x))`` in its steady state, on the two
    layers' own state tensors: y = Q(x * mask), gx = clamp(g) * mask (grad_output clamped in place, like
    DecimalQuantization.backward does, ref quantize.py:65-77, sparse.py backward of x * mask)."""

    @staticmethod
    def forward(ctx, x, xs, layout, mask, param, bits, notch, is_decimal):
        # param: the step's decimal (DecimalQuantizer) or the layer's scale itself (ScalerQuantizer)
        ctx.layout, ctx.mask, ctx.param, ctx.bits, ctx.notch = layout, mask, param, bits, notch
        ctx.is_decimal = is_decimal
        if is_decimal:
            return ops.fq_pow2_fwd(xs, param, layout, mask=mask).view(x.shape)
        return ops.fq_scaler_fwd(xs, param, layout, mask=mask).view(x.shape)

    @staticmethod
    def backward(ctx, grad_output):
        g = grad_output
        if g.dtype != torch.float32:
            g = g.float()
        if not g.is_contiguous():
            g = g.contiguous()
        _, gx = ops.ste_bwd(g, ctx.param, ctx.is_decimal, ctx.bits, ctx.notch, ctx.layout, mask=ctx.mask,
                            clamp_in_place=True, want_gx=True)
        return (gx,) + (None,) * 7


class FusedPruneQuantSequential(nn.Sequential):
    """``Sequential(Sequential(act, PruneLayer), QuantizeLayer)`` (what two ``convert()`` calls build around
    an activation, ref qsparse/convert.py:214-217) or ``Sequential(PruneLayer, QuantizeLayer)``, with the
    SAME children — module tree and ``state_dict`` keys are untouched — whose forward sends every
    steady-state training step through the fused kernels (reduce -> one parameter kernel -> quantize with
    the mask folded in: 20 B/elem with the backward, instead of ~60 + 16).

    A step is fused only when it is provably the plain case: both layers initialised, training, structured
    channel prune (``dimensions={1}``) by the stock ``MagnitudePruningCallback`` (no gradient / l0 / hook
    variants), pruning started and the sparsity not changing at this step, per-tensor stock
    ``DecimalQuantizer`` or ``ScalerQuantizer`` (``quantize()``'s default callback) past its timeout.  Every other step (warm-up, ramp steps, eval, other callbacks,
    non-contiguous inputs) runs the two layers one after the other, exactly as before.  Counters and
    parameters advance identically either way, so the two routes can alternate freely."""

    fused_steps = 0  # how many forwards took the fused route (per instance once incremented)

    def _layers(self):
        first, q = self[0], self[1]
        if isinstance(first, nn.Sequential):
            return first[0], first[1], q
        return None, first, q

    def _fusable(self, p, q, x):
        from .quantize import DecimalQuantizer, QuantizeLayer, ScalerQuantizer
        from .sparse import MagnitudePruningCallback, PruneLayer

        if not (isinstance(p, PruneLayer) and isinstance(q, QuantizeLayer) and self.training and p.training
                and q.training):
            return False
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() >= 2 and x.is_contiguous()):
            return False
        cb, qcb = p.callback, q.callback
        if type(cb) is not MagnitudePruningCallback or type(qcb) not in (DecimalQuantizer, ScalerQuantizer):
            return False
        if cb.use_gradient or cb.l0 or cb.forward_hook is not None or not cb.training or not qcb.training:
            return False
        if p.dimensions != {1} or not p.initted or not q.initted or p.mask.numel() == 1:
            return False
        if p.mask.numel() != x.shape[1] or p.mask.numel() > 2048:
            return False
        n = p._n_mirror.get(p._n_updates)
        if n < p.start or n in p.schedules:
            return False
        if not cb.initted or cb._t() >= cb.stop_mask_refresh or cb.mask_refresh_interval <= 0:
            return False
        if cb.running_average and not hasattr(cb, "magnitude"):
            return False
        if q.channelwise >= 0 or q.timeout <= 0 or q._t_mirror.get(q._n_updates) < q.timeout:
            return False
        if qcb.backward_passthrough or qcb.use_uint or qcb.group_num > 0:
            return False
        if qcb.use_float_scaler != (type(qcb) is ScalerQuantizer):
            return False
        if tuple(q.weight.shape) != (1, 1) or p._s_mirror.get(p._cur_sparsity) < 0:
            return False
        return True

    def _fused_step(self, p, q, x):
        cb, qcb = p.callback, q.callback
        layout = N.channel_layout(x.shape, 1)
        outer, ch, inner = layout
        xs = N.as_f32_contiguous(x.detach())
        sparsity = p._s_mirror.get(p._cur_sparsity)
        t = cb._t()
        refresh = (t % cb.mask_refresh_interval == 0 and t <= cb.stop_mask_refresh) and \
            (t > 0 or not cb.running_average)
        k = kth_rank(sparsity, ch)
        if refresh and k >= ch:    # the un-fused route only indexes sorted()[k] on a refresh step
            raise IndexError(f"index {k} is out of bounds for dimension 0 with size {ch}")
        if not refresh:
            k = min(k, ch - 1)
        is_decimal = not qcb.use_float_scaler
        # DecimalQuantizer: this step's decimal (a fresh tensor: the backward must see THIS step's value);
        # ScalerQuantizer: the quantize kernels read the layer's scale itself, like the reference's saved tensor
        decimal = torch.empty(1, dtype=torch.float32, device=x.device) if is_decimal else None
        with torch.no_grad():
            if cb.running_average:
                magnitude, mode = cb.magnitude.data.view(-1), 1
            else:
                magnitude, mode = torch.empty(ch, dtype=torch.float32, device=x.device), 2
            if graphs.active():
                # graph mode: the step index is the prune callback's own `t` Parameter, read and advanced by the
                # kernel; the quantizer's EMA index keeps its constant distance to it
                if ch > ops.FUSED_STEP_MAX_CHANNELS or cb.stop_mask_refresh != float("inf"):
                    raise graphs.NotCapturable("a fused site with more than "
                                               f"{ops.FUSED_STEP_MAX_CHANNELS} channels or a stopping mask refresh")
                ops.reduce_prune_quant_step(xs, layout, magnitude, p.mask.data.view(-1), q.weight.data.view(-1),
                                            decimal, float(outer * inner), 0, mode, cb.mask_refresh_interval,
                                            kth_rank(sparsity, ch), q.bits, qcb.t - t, True,
                                            step_counter=graphs.callback_counter(cb, x.device))
            elif ch <= ops.FUSED_STEP_MAX_CHANNELS:
                ops.reduce_prune_quant_step(xs, layout, magnitude, p.mask.data.view(-1), q.weight.data.view(-1),
                                            decimal, float(outer * inner), t, mode, refresh, k, q.bits, qcb.t, True)
                cb.t.data.add_(1)
            else:
                ws = ops.reduce_partials(xs, layout)
                ops.prune_quant_step_params(magnitude, p.mask.data.view(-1), q.weight.data.view(-1), decimal, ws,
                                            layout, float(outer * inner), t, mode, refresh, k, q.bits, qcb.t, True)
                cb.t.data.add_(1)
            # the counters of the two layers and their callbacks, as their own forwards advance them
            cb._t_mirror.wrote(cb.t, t + 1)
            n = p._n_mirror.get(p._n_updates)
            p._n_updates.data.add_(1)
            p._n_mirror.wrote(p._n_updates, n + 1)
            qcb.t += 1
            q._quantized = True
            tq = q._t_mirror.get(q._n_updates)
            q._n_updates.data.add_(1)
            q._t_mirror.wrote(q._n_updates, tq + 1)
        self.fused_steps += 1
        return _FusedLayersFn.apply(x, xs, layout, p.mask.data.view(-1),
                                    decimal if is_decimal else q.weight.data.view(-1), q.bits,
                                    1 if qcb.flip_axis else 0, is_decimal)

    def forward(self, x):
        pre, p, q = self._layers()
        if pre is not None:
            x = pre(x)
        if self._fusable(p, q, x):
            return self._fused_step(p, q, x)
        return q(p(x))


# ----------------------------------------------------------------------------- weight chain quantize(prune(layer))
class _FusedWeightFn(torch.autograd.Function):
    """``quantize_layer(prune_layer(w))`` of a weight whose prune mask is frozen (ref qsparse/imitation.py:61-71,
    sparse.py:116, quantize.py:501-508): estimate the row parameters on ``w * mask``, EMA, fake-quantize — one
    launch, 9 B/elem (K8 with an element mask) instead of mask-apply (9) + estimate/quantize (8).  Backward:
    ``clamp(g) * mask`` (Decimal / Scaler, one launch) or ``g * mask`` (Adaptive: identity STE)."""

    @staticmethod
    def forward(ctx, w, ws, mask, q, t_line):
        from .quantize import AdaptiveQuantizer
        qcb = q.callback
        ctx.mask, ctx.rows, ctx.bits = mask, ws.shape[0], q.bits
        counter = graphs.quantizer_counter(qcb, ws.device) if graphs.active() else None
        if type(qcb) is AdaptiveQuantizer:
            if counter is not None:          # graph mode: this call's number is *counter + 1
                y, _ = ops.row_quant_fused_(ws, q.weight.data, ops.ROW_LINE, q.bits, 1, qcb.training, mask=mask,
                                            t_dev=counter)
                counter.add_(1)
            else:
                y, _ = ops.row_quant_fused_(ws, q.weight.data, ops.ROW_LINE, q.bits, t_line, qcb.training, mask=mask)
            ctx.kind = "line"
        else:
            is_decimal = not qcb.use_float_scaler
            kind = ops.ROW_DECIMAL if is_decimal else ops.ROW_SCALER
            if counter is not None:
                y, dec = ops.row_quant_fused_(ws, q.weight.data, kind, q.bits, 0, mask=mask, t_dev=counter)
                counter.add_(1)
            else:
                y, dec = ops.row_quant_fused_(ws, q.weight.data, kind, q.bits, qcb.t, mask=mask)
            qcb.t += 1
            ctx.kind = "ste"
            ctx.is_decimal = is_decimal
            ctx.param = dec if is_decimal else q.weight.data.view(-1)
            ctx.notch = 1 if qcb.flip_axis else 0
            ctx.passthrough = qcb.backward_passthrough
        return y.view(w.shape)

    @staticmethod
    def backward(ctx, grad_output):
        g = N.as_f32_contiguous(grad_output)
        n = g.numel()
        if ctx.kind == "line" or ctx.passthrough:
            return (ops.mask_apply(g, ctx.mask, (1, 1, n)),) + (None,) * 4
        _, gx = ops.ste_bwd(g, ctx.param, ctx.is_decimal, ctx.bits, ctx.notch, (1, ctx.rows, n // ctx.rows),
                            mask=ctx.mask, clamp_in_place=True, want_gx=True)
        return (gx,) + (None,) * 4


def _fused_weight_step(p, q, raw):
    """The fused weight access, or None when this step is not provably the plain frozen-mask case."""
    from .quantize import AdaptiveQuantizer, DecimalQuantizer, QuantizeLayer, ScalerQuantizer
    from .sparse import MagnitudePruningCallback, PruneLayer

    if not (isinstance(p, PruneLayer) and isinstance(q, QuantizeLayer) and p.training and q.training):
        return None
    if not (isinstance(raw, torch.Tensor) and raw.is_cuda and raw.dim() >= 2 and raw.dtype == torch.float32
            and raw.is_contiguous()):
        return None
    cb, qcb = p.callback, q.callback
    if type(cb) is not MagnitudePruningCallback or type(qcb) not in (DecimalQuantizer, ScalerQuantizer,
                                                                     AdaptiveQuantizer):
        return None
    if not p.initted or not q.initted or tuple(p.mask.shape) != tuple(raw.shape):
        return None                                            # unstructured (full-size) masks only
    if cb.use_gradient or cb.forward_hook is not None or not cb.training or not qcb.training or not cb.initted:
        return None
    n = p._n_mirror.get(p._n_updates)
    if n < p.start or n in p.schedules or cb._t() <= cb.stop_mask_refresh:
        return None                                            # the callback still updates magnitude / mask
    if q.channelwise != 0 or q.batch_dimension == 0 or q.timeout <= 0 or qcb.group_num > 0:
        return None
    tq = q._t_mirror.get(q._n_updates)
    if tq < q.timeout:
        return None
    if tuple(q.weight.shape) != (raw.shape[0], qcb.weight_size) or not q.weight.is_contiguous():
        return None
    if type(qcb) is AdaptiveQuantizer and isinstance(qcb.t, torch.Tensor) is False and qcb.t == 0:
        return None                                            # its first optimize() re-creates `t`: unfused
    ws = raw.detach()
    if not ops.row_quant_supported(ws, raw.shape[0]) or p.mask.data_ptr() % 8:
        return None
    with torch.no_grad():
        # the counters of the two layers and their callbacks, as their own forwards advance them
        t = cb._t()
        cb.t.data.add_(1)
        cb._t_mirror.wrote(cb.t, t + 1)
        p._n_updates.data.add_(1)
        p._n_mirror.wrote(p._n_updates, n + 1)
        if graphs.active():
            graphs.quantizer_counter(qcb, raw.device)          # (its device twin exists before the count moves)
        t_line = qcb._next_t() if type(qcb) is AdaptiveQuantizer else 0
        q._quantized = True
        q._n_updates.data.add_(1)
        q._t_mirror.wrote(q._n_updates, tq + 1)
    return _FusedWeightFn.apply(raw, ws, p.mask.data, q, t_line)


def _imitation_chain(mod) -> list:
    """operator names of an ``imitate()`` chain, outermost first (['quantize', 'prune'] = quantize(prune(layer)))"""
    return [c.__dict__["_qsb_imitates"] for c in type(mod).__mro__ if "_qsb_imitates" in c.__dict__]


def _fuse_weight_chain(mod: nn.Module) -> bool:
    if _imitation_chain(mod)[:2] != ["quantize", "prune"] or getattr(type(mod), "_qsb_fused_chain", False):
        return False
    cls = type(mod)

    class FusedChain(cls):
        _qsb_fused_chain = True
        fused_weight_steps = 0

        @property
        def weight(self):
            out = _fused_weight_step(self.prune, self.quantize, self._parameters["weight"])
            if out is not None:
                self.fused_weight_steps += 1
                return out
            return cls.weight.__get__(self)

    FusedChain.__name__ = cls.__name__
    FusedChain.__qualname__ = cls.__qualname__
    mod.__class__ = FusedChain
    return True


def fuse_prune_quantize(model: nn.Module) -> nn.Module:
    """Fusion pass over a converted model (in place): every ``Sequential(Sequential(m, PruneLayer),
    QuantizeLayer)`` / ``Sequential(PruneLayer, QuantizeLayer)`` becomes a ``FusedPruneQuantSequential``
    and every ``quantize(prune(layer))`` weight chain gets a fused ``weight`` property (same children, same
    ``state_dict``).  Returns the model."""
    from .quantize import QuantizeLayer
    from .sparse import PruneLayer

    def is_site(m):
        if type(m) is not nn.Sequential or len(m) != 2 or not isinstance(m[1], QuantizeLayer):
            return False
        first = m[0]
        if isinstance(first, PruneLayer):
            return True
        return type(first) is nn.Sequential and len(first) == 2 and isinstance(first[1], PruneLayer)

    for mod in list(model.modules()):
        if is_site(mod):
            mod.__class__ = FusedPruneQuantSequential
        elif hasattr(mod, "prune") and hasattr(mod, "quantize"):
            _fuse_weight_chain(mod)          # quantize(prune(layer)): frozen-mask steps through K8 with the mask
    return model
