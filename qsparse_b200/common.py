"""Type aliases and tiny helpers shared by the host layer (ref qsparse/common.py)."""
from typing import Union

import torch

TensorOrInt = Union[int, torch.Tensor]
TensorOrFloat = Union[float, torch.Tensor]


def ensure_tensor(v) -> torch.Tensor:
    """Wrap a Python number into a tensor; tensors pass through (ref qsparse/common.py:13-18)."""
    return v if isinstance(v, torch.Tensor) else torch.tensor(v)
