"""Type aliases (reference: lib/typing.py)."""
from typing import Union
import torch

Array = torch.Tensor
ScalarBool = Union[bool, torch.Tensor]
ScalarInt = Union[int, torch.Tensor]
ScalarFloat = Union[float, torch.Tensor]
KeyArray = Union[int, None]
