"""Weight/bias property patching (ref qsparse/imitation.py).

``imitate(module, "quantize", op)`` makes ``module.weight`` return ``op(raw weight)``
each time it is read, by re-classing the instance with a subclass that overrides
the ``weight`` / ``bias`` attributes with properties.  Stacking two calls chains the
operators (``quantize(prune(conv))`` -> prune, then quantize), which is the fusion
site of the prune->quantize kernel.
"""
from typing import Callable, Optional

import torch.nn as nn


def _identity(t):
    return t


def imitate(human: nn.Module, name: str, thing: Callable, bias_thing: Optional[Callable] = None) -> nn.Module:
    base = human.__class__

    def read(self, attr: str):
        # an earlier imitation already made `attr` a property on the class: go through
        # it (that is what chains the operators); otherwise read the raw Parameter.
        descriptor = getattr(base, attr, None)
        if descriptor is not None:
            return descriptor.__get__(self)
        return self._parameters[attr]

    setattr(human, name, thing)
    setattr(human, name + "_bias", bias_thing if bias_thing is not None else _identity)

    class Imitation(base):
        @property
        def weight(self):
            return getattr(self, name)(read(self, "weight"))

        @property
        def bias(self):
            return getattr(self, name + "_bias")(read(self, "bias"))

    Imitation._qsb_imitates = name      # which operator this level of the chain applies (fusion pass)
    Imitation.__name__ = base.__name__
    Imitation.__qualname__ = base.__qualname__
    human.__class__ = Imitation
    return human
