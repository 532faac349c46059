"""The loss of the reference's training loop on the CUDA library: ``WeightedMSELoss`` and
``reshape_tensors_4_loss`` with the reference's signatures and semantics (``modules/loss.py:30-53, 118-160``).

SURVEY.md section 8f rank 2 (the code either side of ``model(X)``).  CUDA fp32 tensors only — like the rest of
the package there is no CPU path.
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from .functional import _require_cuda_f32, _require_same_device, _stream_ptr, _workspace


class _WeightedMSEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, label, weights, reduction: str):
        _require_cuda_f32(pred, "pred")
        _require_cuda_f32(label, "label")
        _require_same_device(pred, label=label, weights=weights)
        lib = _lib.load()
        p, l = pred.contiguous(), label.contiguous()
        B, V, F = p.shape
        wptr = weights.data_ptr() if weights is not None else None
        ctx.reduction = reduction
        ctx.dims = (B, V, F)
        with torch.cuda.device(p.device):
            if reduction == "none":
                out = torch.empty_like(p)
                rc = lib.dsw_wmse_none_fwd(p.data_ptr(), l.data_ptr(), wptr, out.data_ptr(), B, V, F, _stream_ptr(p.device))
                _lib.check(rc, "dsw_wmse_none_fwd")
                ctx.save_for_backward(p, l, *([weights] if weights is not None else []))
                return out
            loss = torch.empty((), dtype=torch.float32, device=p.device)
            ws = _workspace(lib.dsw_wmse_workspace_bytes(), p.device)
            rc = lib.dsw_wmse_fwd(p.data_ptr(), l.data_ptr(), wptr, loss.data_ptr(), ws.data_ptr(), ws.numel(), B, V, F,
                                  1 if reduction == "sum" else 0, _stream_ptr(p.device))
            _lib.check(rc, "dsw_wmse_fwd")
        ctx.save_for_backward(p, l, ws, *([weights] if weights is not None else []))
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        B, V, F = ctx.dims
        saved = ctx.saved_tensors
        g = g.contiguous()
        if not (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            return None, None, None, None
        # d/d label of w (pred - label)^2 is minus d/d pred: the reference's MSE propagates it when the label needs it
        both = lambda grad: (grad if ctx.needs_input_grad[0] else None, -grad if ctx.needs_input_grad[1] else None, None, None)
        if ctx.reduction == "none":
            p, l, *w = saved
            grad = torch.empty_like(p)
            with torch.cuda.device(p.device):
                rc = lib.dsw_wmse_none_bwd(p.data_ptr(), l.data_ptr(), w[0].data_ptr() if w else None, g.data_ptr(), grad.data_ptr(),
                                           B, V, F, _stream_ptr(p.device))
            _lib.check(rc, "dsw_wmse_none_bwd")
            return both(grad)
        p, l, ws, *w = saved
        grad = torch.empty_like(p)
        with torch.cuda.device(p.device):
            rc = lib.dsw_wmse_bwd(p.data_ptr(), l.data_ptr(), w[0].data_ptr() if w else None, ws.data_ptr(), g.data_ptr(),
                                  grad.data_ptr(), B, V, F, _stream_ptr(p.device))
        _lib.check(rc, "dsw_wmse_bwd")
        return both(grad)


class WeightedMSELoss(torch.nn.MSELoss):
    """Area-weighted MSE (reference ``modules/loss.py:118-160``): ``forward(pred, label)`` on ``[batch, node, value]``
    tensors, ``weights`` a 1-D tensor over the nodes (or None), ``reduction`` in ``mean | sum | none`` with the
    reference's normalisations.  One fused kernel per direction."""

    def __init__(self, reduction="mean", weights=None):
        super().__init__(reduction="none")
        if not isinstance(reduction, str) or reduction not in ("mean", "sum", "none"):
            raise ValueError("{} is not a valid value for reduction".format(reduction))
        self.weighted_mse_reduction = reduction
        if weights is not None:
            self.check_weights(weights)
        self.weights = weights

    def forward(self, pred, label):
        if pred.shape != label.shape or pred.dim() != 3:
            raise ValueError(f"pred and label must both be [batch, node, value]; got {tuple(pred.shape)} and {tuple(label.shape)}")
        weights = self.weights
        num_nodes = pred.shape[1]
        if weights is not None:
            if num_nodes != len(weights):  # same check and message as loss.py:135-140
                raise ValueError(
                    "The number of weights does not match the the number of pixels. {} != {}".format(len(weights), num_nodes))
            if weights.device != pred.device or weights.dtype != torch.float32 or not weights.is_contiguous():
                weights = weights.to(device=pred.device, dtype=torch.float32).contiguous()
                self.weights = weights  # (the reference moves them on every call, loss.py:141)
        return _WeightedMSEFunction.apply(pred, label, weights, self.weighted_mse_reduction)

    def check_weights(self, weights):
        if not isinstance(weights, torch.Tensor):
            raise TypeError("Weights type is not a torch.Tensor. Got {}".format(type(weights)))
        if len(weights.shape) != 1:
            raise ValueError("Weights is a 1D vector. Got {}".format(weights.shape))


def reshape_tensors_4_loss(Y_pred, Y_obs, dim_info_dynamic):
    """``[..] -> (data_points, node, feature)`` for both tensors: every dimension other than ``node`` / ``feature``
    is flattened, in its original order (reference ``modules/loss.py:30-53``).  Views where the layout allows."""
    names = [k for k, _ in sorted(dim_info_dynamic.items(), key=lambda item: item[1])]
    keep = [n for n in names if n not in ("node", "feature")]
    perm = [names.index(n) for n in keep] + [names.index("node"), names.index("feature")]

    def one(y):
        yp = y.permute(*perm)
        return yp.reshape(-1, yp.shape[-2], yp.shape[-1])

    return one(Y_pred), one(Y_obs)
