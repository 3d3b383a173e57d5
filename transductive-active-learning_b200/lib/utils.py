"""Index / mask helpers (reference: lib/utils.py:44-75)."""
from __future__ import annotations

import torch

from lib._device import as_f64


def noise_covariance_matrix(k: int, noise_std, default: float = 0) -> torch.Tensor:
    """lib/utils.py:44-50: identity(k) * square(noise_std); the default is
    squared too."""
    from lib._device import device
    if noise_std is None:
        noise_std = default
    ns = as_f64(noise_std)
    return torch.eye(k, dtype=torch.float64, device=device()) * torch.square(ns)


def get_indices(X, A) -> torch.Tensor:
    """lib/utils.py:53-56: indices of the points of `A` within `X` (nearest
    point, first minimum wins) — the tal_nearest_index kernel."""
    from talgp.engine import nearest_index
    X = as_f64(X)
    A = as_f64(A).reshape(-1, X.shape[1])
    # fast path: A is a row-aligned view into X's own storage
    if (isinstance(A, torch.Tensor) and A.is_cuda and A.numel() > 0
            and A.untyped_storage().data_ptr() == X.untyped_storage().data_ptr()
            and A.is_contiguous() and X.is_contiguous()):
        d = X.shape[1]
        off = A.storage_offset() - X.storage_offset()
        if off >= 0 and off % d == 0 and off // d + A.shape[0] <= X.shape[0]:
            start = off // d
            return torch.arange(start, start + A.shape[0], dtype=torch.int64, device=X.device)
    return nearest_index(X, A)


def get_mask(X, A) -> torch.Tensor:
    """lib/utils.py:59-62."""
    X = as_f64(X)
    mask = torch.zeros((X.shape[0],), dtype=torch.bool, device=X.device)
    A = as_f64(A)
    if A.shape[0] > 0:
        mask[get_indices(X, A)] = True
    return mask


def is_subset_of(X, A) -> bool:
    """lib/utils.py:65-71.  NB: membership is tested on *flattened scalars*
    (jnp.isin), not on rows.  (MarginalModel.safe_set_is_subset_of uses the
    tal_scalar_subset kernel for the same test on domain subsets.)"""
    X, A = as_f64(X), as_f64(A)
    if X.shape[0] == 0:
        return A.shape[0] == 0
    if A.shape[0] == 0:
        return False
    return bool(torch.isin(X, A).all().item())


def set_to_idx(arr) -> torch.Tensor:
    """lib/utils.py:74-75."""
    return torch.where(arr)[0]


def grad_approx(x, f, h: float = 1e-3) -> torch.Tensor:
    """lib/utils.py:27-41 — central finite differences of f at x."""
    x = as_f64(x).reshape(-1)
    d = x.shape[0]
    eye = torch.eye(d, dtype=torch.float64, device=x.device)
    X = torch.cat((x + h * eye, x - h * eye, x.reshape(1, -1)), dim=0)
    sample = as_f64(f(X)).reshape(-1)
    return (sample[:d] - sample[d: 2 * d]) / (2.0 * h)
