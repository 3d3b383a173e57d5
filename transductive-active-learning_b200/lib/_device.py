"""Small host-side helpers shared by the drop-in classes."""
from __future__ import annotations

import numpy as np
import torch


def device() -> torch.device:
    if not torch.cuda.is_available():
        from talgp._abi import TalError
        raise TalError("CUDA device required: the discrete-GP path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_f64(x, dev=None) -> torch.Tensor:
    """numpy / list / tensor -> contiguous float64 tensor on the device."""
    dev = dev or device()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev).contiguous()


def to_numpy(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)
