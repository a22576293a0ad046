"""Model conversion: inject prune / quantize operators into an ``nn.Module`` tree.

Behavioural mirror of ``qsparse/convert.py`` (mlzxy/qsparse v2.0.1): same signature,
same traversal order, same layer counting / exclusion rules, same wrapping
(``nn.Sequential(layer, op)`` tagged ``_qsparse_conversion`` for activations,
``quantize(layer)`` / ``prune(layer)`` through ``imitate`` for weights).  Pure host-side
graph surgery, no kernels — it is the caller that creates the prune->quantize
adjacency the fused kernel serves.
"""
from __future__ import annotations

import copy
import warnings
from collections import defaultdict
from typing import List, Mapping, Optional, Sequence, Tuple, Type, Union

import torch.nn as nn

from .quantize import QuantizeLayer, quantize
from .sparse import PruneLayer, prune
from .util import auto_name_prune_quantize_layers, logging, nn_module

_INJECTED_ATTRS = ("quantize", "prune", "quantize_bias")


def _fresh_kwargs(kwargs: Mapping) -> Mapping:
    """every converted layer gets its own copy of module-valued arguments (callbacks)"""
    return {k: (copy.deepcopy(v) if isinstance(v, nn.Module) else v) for k, v in kwargs.items()}


def _type_name(m) -> Optional[str]:
    """class name of a layer / layer type, looking through conversion wrappers"""
    if isinstance(m, nn.Sequential):
        for child in m.children():
            if not isinstance(child, (QuantizeLayer, PruneLayer)):
                return _type_name(child)
        return None
    if isinstance(m, nn.Module):
        return m.__class__.__name__
    return m.__name__


def _is_container(m: nn.Module) -> bool:
    if len(m._modules) == 0:
        return False
    # a leaf whose only children are the operators injected by imitate() is still a leaf
    return not any(hasattr(m, attr) for attr in _INJECTED_ATTRS)


def convert(  # noqa: C901
    model: nn.Module,
    operator: Union[PruneLayer, QuantizeLayer],
    inplace: bool = True,
    weight_layers: Sequence[Type[nn.Module]] = [],
    activation_layers: Sequence[Type[nn.Module]] = [],
    input: bool = False,
    log: bool = True,
    excluded_weight_layer_indexes: Sequence[Tuple[Type[nn.Module], Sequence[int]]] = [],
    excluded_activation_layer_indexes: Sequence[Tuple[Type[nn.Module], Sequence[int]]] = [],
    include: Optional[Union[str, List[str]]] = None,
    exclude: Optional[Union[str, List[str]]] = None,
    order: str = "post",
    fuse: bool = False,
) -> nn.Module:
    """ref qsparse/convert.py:21-245 (argument meaning identical).

    ``fuse`` (extension, off by default): after the conversion, run the fusion pass
    ``fused.fuse_prune_quantize`` over the model, so every ``Sequential(Sequential(act, PruneLayer),
    QuantizeLayer)`` site that this and earlier ``convert()`` calls created sends its steady-state training
    steps through the fused kernels.  Module tree, ``state_dict`` keys and results are unchanged."""
    assert isinstance(operator, (PruneLayer, QuantizeLayer)), \
        "`operator` does not belong to (PruneLayer, QuantizeLayer)"
    assert order in ["pre", "post"], "`order` must be either 'pre' or 'post'"

    must_contain = [include] if isinstance(include, str) else list(include or [])
    must_not_contain = [exclude] if isinstance(exclude, str) else list(exclude or [])

    def skipped(path: str) -> bool:
        return any(s in path for s in must_not_contain)

    def selected(path: str) -> bool:
        return all(s in path for s in must_contain)

    if len(weight_layers) + len(activation_layers) == 0:
        warnings.warn("No weight or activation layers specified, nothing will be converted.")

    def say(msg):
        if log:
            logging.info(msg)

    def make_operator(layer: Optional[nn.Module] = None) -> nn.Module:
        if layer is None:
            return copy.deepcopy(operator)
        wrap = quantize if isinstance(operator, QuantizeLayer) else prune
        return wrap(layer, **_fresh_kwargs(operator._kwargs))

    if not inplace:
        model = copy.deepcopy(model)

    def count(mod: nn.Module, layer_types) -> Mapping[str, int]:
        def walk(m: nn.Module, wanted: str, scope: str) -> int:
            total = 0
            for name, child in m.named_children():
                path = f"{scope}.{name}"
                if skipped(path):
                    continue
                if _is_container(child):
                    total += walk(child, wanted, path)
                elif _type_name(child) == wanted and selected(path):
                    total += 1
            return total

        return {_type_name(t): walk(mod, _type_name(t), "") for t in layer_types}

    def exclusion_table(pairs, totals) -> Mapping[str, Sequence[int]]:
        table = defaultdict(list)
        for cls, indexes in pairs:
            key = _type_name(cls)
            table[key] = [i if i >= 0 else i + totals[key] for i in indexes]
        return table

    root = nn_module(model)
    weight_seen = {_type_name(c): 0 for c in weight_layers}
    weight_excluded = exclusion_table(excluded_weight_layer_indexes, count(model, weight_layers))
    act_seen = {_type_name(c): 0 for c in activation_layers}
    act_excluded = exclusion_table(excluded_activation_layer_indexes, count(model, activation_layers))

    op_name = str(operator).lower()
    for junk in ("(", ")", "layer"):
        op_name = op_name.replace(junk, "")
    op_name = f"`{op_name}`"

    def convert_weights(mod: nn.Module, scope: str = "") -> nn.Module:
        replaced = {}
        for name, child in mod.named_children():
            path = f"{scope}.{name}"
            if skipped(path):
                continue
            if _is_container(child):
                convert_weights(child, path)
                continue
            kind = _type_name(child)
            if kind not in weight_seen:
                continue
            if weight_seen[kind] not in weight_excluded[kind] and selected(path):
                say(f"Apply {op_name} on the {path} weight")
                child = make_operator(child)
                replaced[name] = child
            else:
                say(f"Exclude {path} weight")
            weight_seen[_type_name(child)] += 1
        for name, child in replaced.items():
            mod._modules[name] = child
        return mod

    def convert_activations(mod: nn.Module, scope: str = "") -> nn.Module:
        replaced = {}
        for name, child in mod.named_children():
            path = f"{scope}.{name}"
            if skipped(path):
                continue
            if _is_container(child) and not hasattr(child, "_qsparse_conversion"):
                convert_activations(child, path)
                continue
            kind = _type_name(child)
            if kind not in act_seen:
                continue
            if act_seen[kind] not in act_excluded[kind] and selected(path):
                say(f"Apply {op_name} on the {path} activation")
                wrapped = nn.Sequential(child, make_operator()) if order == "post" \
                    else nn.Sequential(make_operator(), child)
                setattr(wrapped, "_qsparse_conversion", True)
                replaced[name] = wrapped
            else:
                say(f"Exclude {path} activation")
            act_seen[kind] += 1
        for name, child in replaced.items():
            mod._modules[name] = child
        return mod

    def convert_tree(tree: nn.Module) -> nn.Module:
        tree = convert_activations(convert_weights(tree))
        return nn.Sequential(make_operator(), tree) if input else tree

    if model == root:
        model = convert_tree(model)
    else:  # nn.DataParallel-style wrapper
        model.module = convert_tree(model.module)
    auto_name_prune_quantize_layers(nn_module(model))
    if fuse:
        from .fused import fuse_prune_quantize
        fuse_prune_quantize(nn_module(model))
    return model
