"""Shared helpers for the test-suite (golden loading, tolerances)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance: outputs within 1e-4 relative (fp32) of the reference path.
REL_TOL = 1e-4


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def coo_from(g, prefix: str, device="cpu") -> torch.Tensor:
    idx = torch.from_numpy(g[f"{prefix}_idx"].astype(np.int64))
    val = torch.from_numpy(g[f"{prefix}_val"])
    shape = tuple(int(s) for s in g[f"{prefix}_shape"])
    return torch.sparse_coo_tensor(idx, val, shape, check_invariants=False).coalesce().to(device)


def rel_err(a, b) -> float:
    """max |a-b| / max |b|  — the relative error the parity bar is stated in."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def rel_l2(a, b) -> float:
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    denom = b.norm().item()
    return (a - b).norm().item() / (denom if denom > 0 else 1.0)


CONV_CASES = ["conv_cfg1", "conv_first_layer", "conv_last_layer", "conv_k1", "conv_k2_nobias", "conv_k6_wide"]
UNET_CASES = [("unet_max_k3", "max", 3, 10), ("unet_interp_k4", "interp", 4, 11), ("unet_maxval_k3", "maxval", 3, 12)]
