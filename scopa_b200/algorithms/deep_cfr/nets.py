"""The parts of the reference's nets.py that its SDCFR uses (src/algorithms/deep_cfr/nets.py:80-101,
:151-235, :296-331): MLPBlock, FlexibleNet in 'mlp' mode, positive_regret_policy, masked_softmax.
The conv2d_mlp mode is dead code upstream (no caller) and is not provided."""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

_ACTS = {"relu": nn.ReLU, "tanh": nn.Tanh, "gelu": nn.GELU, "identity": nn.Identity, "none": nn.Identity}


def masked_softmax(logits: torch.Tensor, mask: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    very_neg = torch.tensor(-1e9, dtype=logits.dtype, device=logits.device)
    masked = torch.where(mask > 0, logits, very_neg)
    probs = torch.softmax(masked, dim=-1)
    z = (probs * mask).sum(dim=-1, keepdim=True).clamp_min(eps)
    return (probs * mask) / z


def positive_regret_policy(adv: torch.Tensor, mask: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """Convert regrets to a normalized policy using regret matching."""
    pos = torch.relu(adv) * mask
    z = pos.sum(dim=-1, keepdim=True).clamp_min(eps)
    return pos / z


class MLPBlock(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, act: str = "relu", norm: str = "none", dropout: float = 0.0,
                 residual: bool = False):
        super().__init__()
        if norm != "none":
            raise NotImplementedError("the SDCFR advantage net uses mlp_norm='none'")
        self.fc = nn.Linear(in_dim, out_dim)
        self.norm = nn.Identity()
        self.act = _ACTS[act]()
        self.drop = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        self.residual = residual and (in_dim == out_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.drop(self.act(self.norm(self.fc(x))))
        return y + x if self.residual else y


class FlexibleNet(nn.Module):
    def __init__(self, input_shape: Tuple[int, ...], output_dim: int, mode: str = "mlp",
                 mlp_hidden: Optional[List[int]] = None, mlp_act: str = "relu", mlp_norm: str = "none",
                 mlp_dropout: float = 0.0, mlp_residual: bool = False, **unused):
        super().__init__()
        if mode != "mlp":
            raise NotImplementedError("only FlexibleNet(mode='mlp') is used by the reference's SDCFR")
        assert len(input_shape) == 1, "For 'mlp', input_shape must be (D,)."
        self.mode = mode
        layers, last = [], input_shape[0]
        for h in (mlp_hidden or []):
            layers.append(MLPBlock(last, h, act=mlp_act, norm=mlp_norm, dropout=mlp_dropout, residual=mlp_residual))
            last = h
        self.backbone = nn.Sequential(*layers)
        self.head = nn.Linear(last, output_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(self.backbone(x))
