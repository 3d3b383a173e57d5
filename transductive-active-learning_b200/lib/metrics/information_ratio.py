"""Information ratios (reference: lib/metrics/information_ratio.py:10-57) over the directed-ITL
kernels: the individual gains are `itl._directed`, the joint gain is one Schur complement on the
fp64 tensor path per GP (`GaussianDistribution.information_gain`)."""
from __future__ import annotations

from typing import List, Optional

import torch

from lib.algorithms import itl
from lib.model.marginal import ROI, MarginalModel


def compute_strong_information_ratio(F_max: List, prev=None) -> torch.Tensor:
    """:10-28 — min_n min_{j<=n} F_max[j] / F_max[n], updated greedily."""
    if len(F_max) < 1:
        return torch.tensor(float("inf"), dtype=torch.float64)
    vals = torch.stack([torch.as_tensor(v, dtype=torch.float64).reshape(()).cpu() for v in F_max])
    ratios = vals / vals[-1]
    finite = ratios[~torch.isnan(ratios)]
    new = finite.min() if finite.numel() else torch.tensor(float("nan"), dtype=torch.float64)  # nanmin
    if prev is not None:
        prev = torch.as_tensor(prev, dtype=torch.float64).reshape(()).cpu()
        if not bool(torch.isnan(prev)):
            return torch.minimum(new, prev)
    return new


def compute_information_ratio(model: MarginalModel, roi_idx: Optional[torch.Tensor],
                              sample_region_idx: Optional[torch.Tensor]) -> torch.Tensor:
    """:31-57 — sum_{x in B} I(h(A); y(x) | D_n) / I(h(A); y(B) | D_n) with B the sample region."""
    dev = model._gp.device
    if roi_idx is None:
        roi_idx = torch.arange(model.n, device=dev)
    if sample_region_idx is None:
        sample_region_idx = torch.arange(model.n, device=dev)
    roi_idx = torch.as_tensor(roi_idx, device=dev).to(torch.int64).reshape(-1).contiguous()
    boundary_idx = torch.as_tensor(sample_region_idx, device=dev).to(torch.int64).reshape(-1).contiguous()
    noise_rates = model.noise_rate.at(model.domain)
    joint = torch.zeros((), dtype=torch.float64, device=dev)
    for i in range(model.q):
        joint = joint + model.distr(i).information_gain(roi_idx=roi_idx, obs_idx=boundary_idx,
                                                        obs_noise_std=noise_rates[i][boundary_idx])
    indiv = itl._directed(model=model, roi=ROI(model, roi_idx, int(roi_idx.numel())))
    return indiv[boundary_idx].sum() / joint
