"""Run existing qsparse code unchanged: ``install_as_qsparse()`` registers this package (and its
sub-modules, which carry the reference's module names) under the name ``qsparse`` in ``sys.modules``,
so ``import qsparse``, ``from qsparse.quantize import quantize_with_decimal``,
``from qsparse.sparse import PruneLayer`` ... resolve to the B200 implementation.

    import qsparse_b200.compat as compat
    compat.install_as_qsparse()
    import qsparse                      # -> qsparse_b200
"""
from __future__ import annotations

import importlib
import sys

_SUBMODULES = ("common", "convert", "fuse", "imitation", "quantize", "sparse", "util")


def install_as_qsparse(force: bool = False) -> None:
    if "qsparse" in sys.modules and not force:
        existing = sys.modules["qsparse"]
        if getattr(existing, "__name__", "") != "qsparse_b200":
            raise RuntimeError("a different `qsparse` is already imported; pass force=True to shadow it")
        return
    pkg = importlib.import_module("qsparse_b200")
    sys.modules["qsparse"] = pkg
    for name in _SUBMODULES:
        sys.modules[f"qsparse.{name}"] = importlib.import_module(f"qsparse_b200.{name}")
