"""TEST INFRASTRUCTURE ONLY — ``tests/golden/wmse.npz`` from the UNMODIFIED reference ``modules/loss.py``.

    python -m oracle.make_golden_loss          (build container only: needs /root/reference)

``modules/loss.py`` imports plotting / regridding packages that are not installable here (xarray, cartopy,
matplotlib, xsphere); they are stubbed as empty modules — ``WeightedMSELoss`` and ``reshape_tensors_4_loss``
(loss.py:30-53, 118-160) use none of them.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def load_reference_loss():
    ref_import._install_stubs()
    for name in ("xarray", "cartopy", "cartopy.crs", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if ref_import.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_import.REFERENCE_ROOT)
    return importlib.import_module("modules.loss")


def main():
    ref = load_reference_loss()
    r = np.random.default_rng(5)
    out = {}
    for tag, (B, V, F) in {"a": (3, 192, 2), "b": (2, 768, 5)}.items():
        pred = torch.from_numpy(r.standard_normal((B, V, F)).astype(np.float32)).requires_grad_(True)
        label = torch.from_numpy(r.standard_normal((B, V, F)).astype(np.float32))
        w = torch.from_numpy((r.random(V) + 0.1).astype(np.float32))
        w = w / w.sum()
        out[f"{tag}_pred"], out[f"{tag}_label"], out[f"{tag}_w"] = pred.detach().numpy(), label.numpy(), w.numpy()
        for red in ("mean", "sum", "none"):
            for use_w in (True, False):
                crit = ref.WeightedMSELoss(reduction=red, weights=w if use_w else None)
                pred.grad = None
                val = crit(pred, label)
                key = f"{tag}_{red}_{'w' if use_w else 'u'}"
                out[key] = val.detach().numpy()
                if red != "none":
                    val.backward()
                    out[key + "_grad"] = pred.grad.numpy().copy()
    # reshape_tensors_4_loss on a [sample, time, node, feature] tensor
    y = torch.from_numpy(r.standard_normal((2, 3, 48, 2)).astype(np.float32))
    dim_info = {"sample": 0, "time": 1, "node": 2, "feature": 3}
    yp, yo = ref.reshape_tensors_4_loss(y, y + 1, dim_info)
    out["reshape_in"], out["reshape_out"] = y.numpy(), yp.numpy()
    path = os.path.join(ROOT, "tests", "golden", "wmse.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
