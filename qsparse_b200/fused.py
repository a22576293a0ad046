"""Fused structured prune -> pow2 quantize (K7): the training step of
``Sequential(PruneLayer(dimensions={channel}), QuantizeLayer(channelwise=-1,
callback=DecimalQuantizer()))`` (the adjacency ``convert()`` creates,
ref qsparse/convert.py:214-217) in four launches and 20 B/elem:

    forward   reduce_stats(x)            4 B/elem   per-channel sum|x| and max|x|, one read
              prune_quant_params         ~0         magnitude EMA, k-th threshold, mask,
                                                    abs-max of kept channels, scale EMA, decimal
              fq_pow2_fwd(x, mask)       8 B/elem   y = Q(x * mask); pruned channels are not read
    backward  ste_bwd(g, mask)           8 B/elem   gx = clamp(g) * mask

against the reference's ~60 B/elem forward.  No host synchronisation:
step counters are host integers, every derived scalar stays on the device.  With a
process group the statistics row is all-gathered (parallel.StatExchange) and combined
in rank order inside the parameter kernel.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _native as N
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
                 mask_refresh_interval: int = 1, group=None, mutate_grad_output: bool = False):
        super().__init__()
        self.sparsity = sparsity
        self.bits = bits
        self.channel_index = channel_index
        self.running_average = running_average
        self.mask_refresh_interval = max(int(mask_refresh_interval), 1)
        self.group = group
        self.mutate_grad_output = mutate_grad_output  # also clamp grad_output in place (quantize.py:72)
        self.t_prune = 0   # MagnitudePruningCallback.t
        self.t_quant = 0   # DecimalQuantizer.t
        self._exchange: Optional[StatExchange] = None

    def _layout(self, x):
        return N.channel_layout(x.shape, self.channel_index)

    def _allocate(self, x, channels):
        dev = x.device
        self.magnitude = nn.Parameter(torch.zeros(channels, device=dev), requires_grad=False)
        self.mask = nn.Parameter(torch.ones(channels, dtype=torch.bool, device=dev), requires_grad=False)
        self.scale = nn.Parameter(torch.zeros(1, 1, device=dev), requires_grad=False)
        self.decimal = torch.zeros(1, device=dev)
        self._p2p = make_exchange(channels, dev, self.group) if channels <= 2048 else None
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
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
                # one kernel: finalize + peer-memory exchange + parameters
                ws = ops.reduce_partials(xs, layout)
                world = self._p2p.world if self._p2p is not None else 1
                ops.prune_quant_step_params(self.magnitude.data, self.mask.data, self.scale.data, self.decimal, ws,
                                            layout, float(outer * inner * world), t, mode, refresh, k, self.bits,
                                            self.t_quant, True,
                                            group=self._p2p.handle if self._p2p is not None else None,
                                            step_stamp=self._p2p.next_stamp() if self._p2p is not None else 1)
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
