"""Model without safety constraints (reference: lib/model/unsafe.py:9-13)."""
import torch

from lib.model.marginal import MarginalModel


class UnsafeModel(MarginalModel):
    @property
    def operating_safe_set(self) -> torch.Tensor:
        return torch.ones((self.n,), dtype=torch.bool, device=self.domain.device)
