"""Per-step metrics of the transductive-learning experiments (reference:
lib/metrics/transductive_learning.py:10-37): entropy and standard deviations inside the ROI."""
from __future__ import annotations

import torch

from lib import _random


def transductive_learning_metrics_generator(roi_mask, entropy_jitter=None):
    def transductive_learning_metrics(model, f, x):
        if entropy_jitter is not None:
            key = entropy_jitter["key"]
            entropy_jitter["key"] = int(key) + 1
            jitter = entropy_jitter["amount"] * float(_random.uniform(key, (1,))[0])
        else:
            jitter = 0.0
        mask = torch.as_tensor(roi_mask, device=model.stddevs.device).reshape(-1).to(torch.bool)
        sd = model.stddevs[:, mask]
        return dict(
            entropy_in_roi=model.masked_entropy(mask, jitter=jitter),
            max_stddev_in_roi=sd.max() if sd.numel() else torch.zeros((), dtype=torch.float64, device=sd.device),
            avg_stddev_in_roi=sd.mean(),
        )

    return transductive_learning_metrics
