"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's area-weighted MSE loss
(``modules/loss.py:118-148`` ``WeightedMSELoss.forward``; ``reshape_tensors_4_loss`` ``:30-53``).

Plain torch on the CPU; pinned against vectors produced by the unmodified reference class
(``oracle/make_golden_loss.py`` -> ``tests/golden/wmse.npz``).  Nothing in the product imports this file.
"""
from __future__ import annotations

import torch


def weighted_mse(pred: torch.Tensor, label: torch.Tensor, weights: torch.Tensor | None, reduction: str = "mean"):
    """``pred``, ``label``: ``[batch, node, value]``; ``weights``: ``[node]`` or None (loss.py:129-148)."""
    mse = (pred - label) ** 2                                    # nn.MSELoss(reduction="none"), loss.py:119
    n_batch, num_nodes, n_val = mse.shape
    if weights is None:
        weights = torch.ones(num_nodes, dtype=mse.dtype)          # loss.py:133-134
    if num_nodes != len(weights):                                 # loss.py:135-140
        raise ValueError(
            "The number of weights does not match the the number of pixels. {} != {}".format(len(weights), num_nodes))
    weights = weights.view(1, -1, 1)                              # loss.py:141 (rebinds `weights`)
    weighted = mse * weights                                      # loss.py:142
    if reduction == "sum":
        # loss.py:143-144 multiplies by len(weights) AFTER the view to [1, V, 1], i.e. by 1 — pinned by the
        # golden vectors made from the unmodified class
        return torch.sum(weighted) * len(weights)
    if reduction == "mean":
        return torch.sum(weighted) / torch.sum(weights) / n_batch / n_val   # loss.py:145-146
    return weighted                                               # loss.py:147-148


def reshape_4_loss(y: torch.Tensor, dim_order):
    """``[..dims..] -> [data_points, node, feature]`` with every dimension other than node / feature flattened in
    their original order (loss.py:30-53)."""
    names = list(dim_order)
    keep = [n for n in names if n not in ("node", "feature")]
    perm = [names.index(n) for n in keep] + [names.index("node"), names.index("feature")]
    yp = y.permute(*perm)
    return yp.reshape(-1, yp.shape[-2], yp.shape[-1])
