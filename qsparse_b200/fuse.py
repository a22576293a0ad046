"""Fold batch-norm layers into the convolution / linear layer in front of them
(ref qsparse/fuse.py:25-163): an offline, one-shot rewrite of a model's weights before
deployment — plain tensor algebra on a few parameters, no hot path, no kernels of ours.

    w' = w * gamma / sqrt(var + 1e-5)      (along the layer's output-channel axis)
    b' = (b - mean) * gamma / sqrt(var + 1e-5) + beta

Behaviour kept from the reference: the epsilon is the constant 1e-5 (not ``bn.eps``); a layer
without bias gains one; ``nn.Sequential`` containers are rebuilt without the fused batch-norms,
a container left with one child is replaced by that child, an empty one disappears; a
batch-norm that opens a nested ``Sequential`` is folded into the layer that precedes the
container; only ``Sequential`` children of a non-``Sequential`` root are visited.
"""
from __future__ import annotations

import copy
from typing import Callable, Dict, Iterable, Mapping, Optional, Tuple

import torch
import torch.nn as nn

from .util import logging, nn_module

BNFuser = Callable[[nn.Module, nn.Module], nn.Module]


def _fold(layer: nn.Module, bn: nn.Module, channel_axis: int) -> nn.Module:
    weight = layer._parameters["weight"].detach()
    bias = layer._parameters["bias"].detach() if layer.bias is not None else 0
    inv_std = bn.weight.detach() / torch.sqrt(bn.running_var.detach().add(1e-5))
    shape = [1] * weight.dim()
    shape[channel_axis] = -1
    layer._parameters["weight"].data = weight * inv_std.view(shape)
    layer._parameters["bias"] = nn.Parameter((bias - bn.running_mean.detach()) * inv_std + bn.bias.detach())
    return layer


def conv2d_bn_fuser(conv: nn.Module, bn: nn.Module) -> nn.Module:
    """Conv2d weights are [out, in, kh, kw] (ref qsparse/fuse.py:25-38)"""
    return _fold(conv, bn, 0)


def linear_bn_fuser(linear: nn.Module, bn: nn.Module) -> nn.Module:
    """Linear weights are [out, in] (ref qsparse/fuse.py:41-54)"""
    return _fold(linear, bn, 0)


def deconv2d_bn_fuser(deconv: nn.Module, bn: nn.Module) -> nn.Module:
    """ConvTranspose2d weights are [in, out, kh, kw] (ref qsparse/fuse.py:57-70)"""
    return _fold(deconv, bn, 1)


default_handlers: Dict[str, BNFuser] = dict(Conv2d=conv2d_bn_fuser, Linear=linear_bn_fuser,
                                            ConvTranspose2d=deconv2d_bn_fuser)


def _is_batchnorm(m: nn.Module) -> bool:
    return type(m).__name__.lower().startswith("batchnorm")


def fuse_bn(model: nn.Module, layers: Iterable[str] = ("Conv2d", "Linear", "ConvTranspose2d"),
            handlers: Optional[Mapping[str, BNFuser]] = None, log: bool = True, inplace: bool = True) -> nn.Module:
    """ref qsparse/fuse.py:76-163 (same arguments)."""
    table = {**default_handlers, **(handlers or {})}
    wanted = set(layers)
    for name in wanted:
        assert name in table, f"layer {name} is not in handlers"
    if not inplace:
        model = copy.deepcopy(model)

    def rebuild(seq: nn.Sequential, before: Optional[nn.Module]) -> Tuple[Optional[nn.Module], Optional[nn.Module]]:
        """-> (what replaces `seq`, the possibly re-written layer that preceded it)"""
        kept = []

        def previous():
            return kept[-1] if kept else before

        def replace_previous(m):
            nonlocal before
            if kept:
                kept[-1] = m
            else:
                before = m

        for child in seq.children():
            if _is_batchnorm(child):
                target = previous()
                kind = type(target).__name__ if target is not None else ""
                if kind in wanted:
                    if log:
                        logging.info(f"Fuse {child} into {target}")
                    replace_previous(table[kind](target, child))
                else:
                    kept.append(child)
            elif isinstance(child, nn.Sequential):
                inner, rewritten = rebuild(child, previous())
                if rewritten is not None:
                    replace_previous(rewritten)
                if inner is not None:
                    kept.append(inner)
            else:
                kept.append(child)
        if not kept:
            return None, before
        return (kept[0] if len(kept) == 1 else nn.Sequential(*kept)), before

    root = nn_module(model)
    if isinstance(root, nn.Sequential):
        fused = rebuild(root, None)[0]
        if model is root:
            model = fused
        else:
            model.module = fused
    else:
        for name, child in root.named_children():
            if isinstance(child, nn.Sequential):
                root._modules[name] = rebuild(child, None)[0]
    return model
