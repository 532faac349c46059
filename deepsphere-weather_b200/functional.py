"""Autograd functions over the C-ABI (``libdsw.so``) — the host-side half of the hot path.

Each ``torch.autograd.Function`` maps to one row of SURVEY.md §8a.  Tensors cross the boundary as
raw device pointers + element strides; PyTorch only provides memory, streams and autograd
bookkeeping.  Non-CUDA tensors raise (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import torch
from torch.autograd.function import once_differentiable

from . import _lib


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda_f32(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} must be a CUDA tensor: deepsphere_weather_b200 has no CPU path "
            f"(got device {t.device})."
        )
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")


def _channel_last(x: torch.Tensor) -> torch.Tensor:
    """The kernels take arbitrary batch / node strides but need unit feature stride."""
    if x.stride(2) != 1 and x.shape[2] != 1:
        return x.contiguous()
    if x.shape[2] == 1 and x.stride(2) != 1:
        return x.contiguous()
    return x


# Keep the forward's Chebyshev terms for the weight gradient (what the reference's autograd does
# too) instead of recomputing them: one third fewer SpMM hops per step for (K-1)/K of an activation
# of extra memory per layer.  DSW_SAVE_TERMS=0 (or set_save_terms(False)) recomputes.
_SAVE_TERMS = os.environ.get("DSW_SAVE_TERMS", "1") != "0"
_OPT_L2_CHUNK = 1


def set_save_terms(flag: bool) -> None:
    global _SAVE_TERMS
    _SAVE_TERMS = bool(flag)


def _require_same_device(ref: torch.Tensor, **others):
    """The kernels take raw pointers: tensors of different devices would be a peer access fault or a silent cross-device
    read where the reference raises a device-mismatch error."""
    for name, t in others.items():
        dev = getattr(t, "device", None)
        if t is not None and dev is not None and dev != ref.device:
            raise RuntimeError(f"{name} is on {dev} but the inputs are on {ref.device}: all operands must share one CUDA device")


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------------
# Plans
# --------------------------------------------------------------------------------------------


# Plans whose Python owner has been garbage-collected.  Their device memory is NOT released from the finaliser: the
# garbage collector may run at any moment — in particular while another part of the program is capturing a CUDA graph,
# where cudaFree (an implicitly synchronising call) invalidates the capture.  They are released the next time a plan is
# built (plan construction synchronises anyway and never happens inside a capture) or by release_dead_plans().
_DEAD_PLANS: list = []


def release_dead_plans() -> int:
    """Free the device memory of plans whose owners are gone; returns how many were released."""
    n = 0
    while _DEAD_PLANS:
        _lib.load().dsw_plan_destroy(_DEAD_PLANS.pop())
        n += 1
    return n


class SparsePlan:
    """Device-resident CSR / CSR^T / row-block layouts of one sparse operator (``dsw_plan``).

    Built from exactly what the reference stores in its module buffers: a coalesced
    ``torch.sparse_coo_tensor`` with int64 indices and fp32 values (``layers.py:82-106,584-594``).
    """

    def __init__(self, coo: torch.Tensor):
        if not coo.is_sparse:
            raise TypeError("SparsePlan expects a torch sparse COO tensor")
        if not coo.is_cuda:
            raise RuntimeError("SparsePlan needs the sparse operator on a CUDA device (no CPU path)")
        coo = coo.coalesce()
        idx = coo.indices().contiguous()
        val = coo.values().to(torch.float32).contiguous()
        self.shape = tuple(coo.shape)
        self.nnz = int(val.numel())
        self.device = coo.device
        lib = _lib.load()
        release_dead_plans()
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = lib.dsw_plan_create(
                self.shape[0], self.shape[1], self.nnz, idx[0].data_ptr(), idx[1].data_ptr(), val.data_ptr(),
                _stream_ptr(self.device), C.byref(handle),
            )
        _lib.check(rc, "dsw_plan_create")
        self.handle = handle
        self._finalizer = weakref.finalize(self, _DEAD_PLANS.append, handle)

    @property
    def operand_bytes(self) -> int:
        return int(_lib.load().dsw_plan_operand_bytes(self.handle))


_PLAN_CACHE: dict = {}        # content digest -> SparsePlan (strong; bounded, least recently used first out)
_PLAN_BY_TENSOR: dict = {}    # id(tensor) -> (weakref to the tensor, its _version, plan); dropped when the tensor dies
_PLAN_CACHE_MAX = 64


def _operator_digest(c: torch.Tensor) -> tuple:
    """Content key of a coalesced sparse COO operator: a 128-bit digest of the exact (indices, values) bytes.  (A
    checksum of sums is not enough: all permutation matrices of one size, or two one-hot pooling matrices whose winner
    indices add up alike, would share it and silently reuse each other's plan.)"""
    import hashlib

    h = hashlib.blake2b(digest_size=16)
    h.update(c.indices().cpu().contiguous().numpy().tobytes())
    h.update(c.values().to(torch.float32).cpu().contiguous().numpy().tobytes())
    return (c.device.index, tuple(c.shape), int(c.values().numel()), h.hexdigest())


def plan_for(coo: torch.Tensor) -> SparsePlan:
    """Plan cache keyed by operator content (modules of one U-Net level share a Laplacian but ``module.to(device)`` gives
    each its own copy of the buffer).  The digest costs one device-to-host copy per *new* buffer object; afterwards the
    buffer is recognised by identity and version."""
    if not coo.is_cuda:
        raise RuntimeError("sparse operator must live on a CUDA device (no CPU path)")
    ident = id(coo)
    hit = _PLAN_BY_TENSOR.get(ident)
    if hit is not None and hit[0]() is coo and hit[1] == coo._version:
        return hit[2]
    c = coo.coalesce()
    key = _operator_digest(c)
    plan = _PLAN_CACHE.pop(key, None)
    if plan is None:
        plan = SparsePlan(c)
    _PLAN_CACHE[key] = plan  # (re-)inserted last: most recently used
    while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
        _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
    _PLAN_BY_TENSOR[ident] = (weakref.ref(coo, lambda _r, ident=ident: _PLAN_BY_TENSOR.pop(ident, None)), coo._version, plan)
    return plan


# --------------------------------------------------------------------------------------------
# Chebyshev convolution                                           reference layers.py:113-180, 365-376
# --------------------------------------------------------------------------------------------


class ChebConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, plan: SparsePlan, act: int = 0, input_is_relu: bool = False, premasked: bool = False):
        """``input_is_relu``: x is the output of a ReLU whose backward the caller has delegated to this layer — the
        backward returns dx * [x > 0] (mask applied by the kernel that writes dx).  ``premasked``: this layer's own fused
        ReLU (act = 1) gets its gradient already masked by its only consumer, so the backward neither masks nor keeps y."""
        _require_cuda_f32(x, "inputs")
        _require_cuda_f32(weight, "weight")
        B, V, Fin = x.shape
        Fin_w, K, Fout = weight.shape
        if Fin != Fin_w:  # same check and message as layers.py:149-154
            raise ValueError(
                "Input tensor shape does not match the expected shape: \n"
                + "- Input tensor shape :{} \n".format(Fin)
                + "- Expected tensor shape :{} \n".format(Fin_w)
            )
        if V != plan.shape[0]:
            raise ValueError(f"inputs have {V} nodes but the laplacian is {plan.shape}")
        _require_same_device(x, weight=weight, bias=bias, laplacian=plan)
        lib = _lib.load()
        x = _channel_last(x)
        w = weight.contiguous()
        bptr = None
        if bias is not None:
            _require_cuda_f32(bias, "bias")
            bptr = bias.contiguous().data_ptr()
        y = torch.empty((B, V, Fout), dtype=torch.float32, device=x.device)
        ws = _workspace(lib.dsw_cheb_fwd_workspace_bytes(B, V, Fin, Fout, K), x.device)
        with torch.cuda.device(x.device):
            rc = lib.dsw_cheb_fwd(
                plan.handle, x.data_ptr(), x.stride(0), x.stride(1), w.data_ptr(), bptr, y.data_ptr(),
                B, Fin, Fout, K, int(act), ws.data_ptr(), ws.numel(), _stream_ptr(x.device),
            )
        _lib.check(rc, "dsw_cheb_fwd")
        # Only x and W are saved: callers modify our output in place (`x_out *= rezero_weight`,
        # my_models_graph.py:213), so the backward recomputes the Chebyshev terms instead.
        # The forward's terms of x are kept for the weight gradient only when both directions use the
        # orders that produce / consume them (dsw_cheb.cu); otherwise dW comes from the terms of dy.
        keep = (_SAVE_TERMS and K > 1 and lib.dsw_get_option(_OPT_L2_CHUNK) <= 1
                and lib.dsw_cheb_fwd_algo(Fin, Fout, K) == 1 and lib.dsw_cheb_bwd_algo(Fin, Fout, K) == 2
                and (ctx.needs_input_grad[1] or (bias is not None and ctx.needs_input_grad[2])))
        # act = 1 (ReLU fused into the last kernel of the forward): the backward masks dy with the sign of the
        # output, which is therefore saved — such a layer's output must not be modified in place by the caller
        # (the reference applies its activation out of place, my_models_graph.py:113).
        own_mask = bool(act) and not premasked
        ctx.save_for_backward(x, w, *([ws] if keep else []), *([y] if own_mask else []))
        ctx.plan = plan
        ctx.has_bias = bias is not None
        ctx.act = int(own_mask)
        ctx.keep = bool(keep)
        ctx.mask_dx = bool(input_is_relu)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w, *saved = ctx.saved_tensors
        if ctx.act:
            dy = torch.ops.aten.threshold_backward(dy, saved.pop(), 0.0)  # ReLU mask (what F.relu's backward does)
        terms_ptr = saved[0].data_ptr() if ctx.keep else None
        plan = ctx.plan
        lib = _lib.load()
        B, V, Fin = x.shape
        _, K, Fout = w.shape
        dy = dy.contiguous()
        need_dx = ctx.needs_input_grad[0]
        need_dw = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        if not (need_dx or need_dw):
            return None, None, None, None, None, None, None
        flags = 0
        post_mask = False
        if ctx.mask_dx and need_dx:
            if B == 1 or x.stride(0) == V * x.stride(1):
                flags = 1  # DSW_BWD_MASK_DX_BY_X
            else:
                post_mask = True
        dx = torch.empty((B, V, Fin), dtype=torch.float32, device=x.device) if need_dx else None
        dw = torch.empty_like(w) if need_dw else None
        db = torch.empty(Fout, dtype=torch.float32, device=x.device) if (need_dw and ctx.has_bias) else None
        have_saved = 1 if (terms_ptr is not None or not need_dw) else 0
        with torch.cuda.device(x.device):
            ws = _workspace(lib.dsw_cheb_bwd_workspace_bytes(B, V, Fin, Fout, K, have_saved), x.device)
            rc = lib.dsw_cheb_bwd_ex(
                plan.handle, x.data_ptr(), x.stride(0), x.stride(1), dy.data_ptr(), w.data_ptr(), terms_ptr,
                dx.data_ptr() if dx is not None else None, dw.data_ptr() if dw is not None else None,
                db.data_ptr() if db is not None else None, B, Fin, Fout, K, flags, ws.data_ptr(), ws.numel(),
                _stream_ptr(x.device),
            )
        _lib.check(rc, "dsw_cheb_bwd_ex")
        if post_mask:
            dx = torch.ops.aten.threshold_backward(dx, x, 0.0)
        return dx, dw, db, None, None, None, None


def cheb_conv(x, weight, bias, plan: SparsePlan, act: int = 0, input_is_relu: bool = False, premasked: bool = False):
    return ChebConvFunction.apply(x, weight, bias, plan, act, input_is_relu, premasked)


def cheb_terms(x: torch.Tensor, plan: SparsePlan, K: int) -> torch.Tensor:
    """The recurrence alone: ``[K-1, B, V, F]`` holding T_1 x .. T_{K-1} x (layers.py:163-169)."""
    _require_cuda_f32(x, "inputs")
    _require_same_device(x, laplacian=plan)
    x = _channel_last(x)
    B, V, F = x.shape
    out = torch.empty((max(K - 1, 0), B, V, F), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().dsw_cheb_terms(
            plan.handle, x.data_ptr(), x.stride(0), x.stride(1), out.data_ptr(), B, F, K, _stream_ptr(x.device)
        )
    _lib.check(rc, "dsw_cheb_terms")
    return out


# --------------------------------------------------------------------------------------------
# Per-node linear map (ResBlock skip connection)          reference my_models_graph.py:196-201, 214
# --------------------------------------------------------------------------------------------


class NodeLinearFunction(torch.autograd.Function):
    """``y = x @ W^T + b`` on ``[B, V, Fin]`` with torch.nn.Linear's parameter layout, through the
    tcgen05 channel-mix / weight-gradient kernels (``dsw_linear_*``)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _require_cuda_f32(x, "x")
        _require_cuda_f32(weight, "weight")
        B, V, Fin = x.shape
        Fout, Fin_w = weight.shape
        if Fin != Fin_w:
            raise ValueError(f"input has {Fin} features but the linear layer expects {Fin_w}")
        lib = _lib.load()
        x = x.contiguous()
        w = weight.contiguous()
        bptr = bias.contiguous().data_ptr() if bias is not None else None
        y = torch.empty((B, V, Fout), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            ws = _workspace(lib.dsw_linear_workspace_bytes(B, V, Fin, Fout), x.device)
            rc = lib.dsw_linear_fwd(x.data_ptr(), x.stride(0), x.stride(1), w.data_ptr(), bptr, y.data_ptr(), B, V, Fin,
                                    Fout, ws.data_ptr(), ws.numel(), _stream_ptr(x.device))
        _lib.check(rc, "dsw_linear_fwd")
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        lib = _lib.load()
        B, V, Fin = x.shape
        Fout = w.shape[0]
        dy = dy.contiguous()
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_db = ctx.has_bias and ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if need_dx else None
        dw = torch.empty_like(w) if need_dw else None
        db = torch.empty(Fout, dtype=torch.float32, device=x.device) if need_db else None
        if need_dx or need_dw or need_db:
            with torch.cuda.device(x.device):
                ws = _workspace(lib.dsw_linear_workspace_bytes(B, V, Fin, Fout), x.device)
                rc = lib.dsw_linear_bwd(
                    x.data_ptr(), x.stride(0), x.stride(1), dy.data_ptr(), w.data_ptr(),
                    dx.data_ptr() if dx is not None else None, dw.data_ptr() if dw is not None else None,
                    db.data_ptr() if db is not None else None, B, V, Fin, Fout, ws.data_ptr(), ws.numel(),
                    _stream_ptr(x.device),
                )
            _lib.check(rc, "dsw_linear_bwd")
        return dx, dw, db


# --------------------------------------------------------------------------------------------
# ResBlock tail: ReZero scale + residual add              reference my_models_graph.py:211-215
# --------------------------------------------------------------------------------------------


class RezeroResidualFunction(torch.autograd.Function):
    """``y = rezero_weight * conv_out + skip`` in one streaming pass (``dsw_rezero_fwd``); the backward
    produces ``rezero_weight * g`` and ``sum(g * conv_out)`` in one pass (``dsw_rezero_bwd``) and hands
    ``g`` itself to the skip branch.  Same arithmetic as the reference's two in-place updates."""

    @staticmethod
    def forward(ctx, conv_out, skip, weight):
        _require_cuda_f32(conv_out, "conv_out")
        _require_cuda_f32(skip, "skip")
        _require_cuda_f32(weight, "rezero_weight")
        if conv_out.shape != skip.shape:
            raise ValueError(f"residual shapes differ: {tuple(conv_out.shape)} vs {tuple(skip.shape)}")
        if weight.numel() != 1:
            raise ValueError("rezero_weight must hold one element")
        a = conv_out.contiguous()
        s = skip.contiguous()
        y = torch.empty_like(a)
        with torch.cuda.device(a.device):
            rc = _lib.load().dsw_rezero_fwd(a.data_ptr(), s.data_ptr(), weight.data_ptr(), y.data_ptr(), a.numel(),
                                            _stream_ptr(a.device))
        _lib.check(rc, "dsw_rezero_fwd")
        ctx.save_for_backward(a, weight)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        a, weight = ctx.saved_tensors
        lib = _lib.load()
        g = g.contiguous()
        need_a, need_s, need_w = ctx.needs_input_grad
        da = torch.empty_like(a) if need_a else None
        dw = torch.empty_like(weight) if need_w else None
        if need_a or need_w:
            with torch.cuda.device(a.device):
                ws = _workspace(lib.dsw_rezero_bwd_workspace_bytes(), a.device)
                rc = lib.dsw_rezero_bwd(g.data_ptr(), a.data_ptr(), weight.data_ptr(),
                                        da.data_ptr() if da is not None else None,
                                        dw.data_ptr() if dw is not None else None, ws.data_ptr(), ws.numel(), a.numel(),
                                        _stream_ptr(a.device))
            _lib.check(rc, "dsw_rezero_bwd")
        return da, (g if need_s else None), dw


def rezero_residual(conv_out, skip, weight):
    return RezeroResidualFunction.apply(conv_out, skip, weight)


class _ForkState:
    """Hand-over between a ResBlock's tail and its ``ForkFunction`` node (see ``fork``)."""

    __slots__ = ("g", "w")

    def __init__(self):
        self.g = None
        self.w = None


class ForkFunction(torch.autograd.Function):
    """Identity on the way in; on the way back it finishes the ResBlock's input gradient.  A ResBlock's input feeds the
    convolution branch and the skip connection (``my_models_graph.py:205-215``), so autograd would add two gradients in a
    separate element-wise pass.  Here the tail (``LinearRezeroFunction``) leaves its output gradient ``g`` in ``state``
    instead of computing the skip's input gradient, and this node — which runs once the convolution branch's gradient has
    arrived — computes ``dx = g . Wl + d_conv`` in the channel-mix epilogue (``dsw_linear_bwd_acc``)."""

    @staticmethod
    def forward(ctx, x, state: _ForkState):
        ctx.state = state
        return x.as_strided(x.shape, x.stride(), x.storage_offset())

    @staticmethod
    @once_differentiable
    def backward(ctx, d):
        st = ctx.state
        g, w = st.g, st.w
        st.g = st.w = None
        if g is None:
            return d, None
        lib = _lib.load()
        B, V, Fin = d.shape
        Fout = w.shape[0]
        d = d.contiguous()
        dx = torch.empty_like(d)
        with torch.cuda.device(d.device):
            ws = _workspace(lib.dsw_linear_workspace_bytes(B, V, Fin, Fout), d.device)
            rc = lib.dsw_linear_bwd_acc(None, 0, 0, g.data_ptr(), w.data_ptr(), d.data_ptr(), dx.data_ptr(), None, None, B, V, Fin,
                                        Fout, ws.data_ptr(), ws.numel(), _stream_ptr(d.device))
        _lib.check(rc, "dsw_linear_bwd_acc")
        return dx, None


def fork(x):
    """``(x_for_the_convolution_branch, state)``; hand ``state`` to ``linear_rezero(..., fork=state)``."""
    state = _ForkState()
    return ForkFunction.apply(x, state), state


class LinearRezeroFunction(torch.autograd.Function):
    """The whole ResBlock tail when the skip connection is a Linear, in one launch:
    ``y = x @ W^T + b + rezero_weight * conv_out`` (``dsw_linear_rezero_fwd``: the channel-mix epilogue adds the
    scaled convolution branch).  Backward = ``dsw_linear_bwd`` with ``dy = g`` and ``dsw_rezero_bwd``.

    ``cat_slot``: the output is written as the SECOND half of a fresh ``[B, V, 2 Fout]`` buffer (a strided view is
    returned) so that the decoder's skip concatenation needs no copy (``unpool_cat``).  ``fork``: see ``ForkFunction``."""

    @staticmethod
    def forward(ctx, x, weight, bias, conv_out, rezero_weight, cat_slot=False, fork_state=None):
        for t, n in ((x, "x"), (weight, "weight"), (conv_out, "conv_out"), (rezero_weight, "rezero_weight")):
            _require_cuda_f32(t, n)
        B, V, Fin = x.shape
        Fout, Fin_w = weight.shape
        if Fin != Fin_w or tuple(conv_out.shape) != (B, V, Fout) or rezero_weight.numel() != 1:
            raise ValueError("shape mismatch in the fused residual tail")
        _require_same_device(x, weight=weight, bias=bias, conv_out=conv_out, rezero_weight=rezero_weight)
        lib = _lib.load()
        x = x.contiguous()
        w = weight.contiguous()
        a = conv_out.contiguous()
        bptr = bias.contiguous().data_ptr() if bias is not None else None
        if cat_slot:
            buf = torch.empty((B, V, 2 * Fout), dtype=torch.float32, device=x.device)
            y = buf[:, :, Fout:]
        else:
            y = torch.empty((B, V, Fout), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            ws = _workspace(lib.dsw_linear_workspace_bytes(B, V, Fin, Fout), x.device)
            rc = lib.dsw_linear_rezero_fwd_ld(x.data_ptr(), x.stride(0), x.stride(1), w.data_ptr(), bptr, a.data_ptr(),
                                              rezero_weight.data_ptr(), y.data_ptr(), y.stride(1), B, V, Fin, Fout, ws.data_ptr(),
                                              ws.numel(), _stream_ptr(x.device))
        _lib.check(rc, "dsw_linear_rezero_fwd_ld")
        ctx.save_for_backward(x, w, a, rezero_weight)
        ctx.has_bias = bias is not None
        ctx.fork_state = fork_state
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w, a, rz = ctx.saved_tensors
        lib = _lib.load()
        B, V, Fin = x.shape
        Fout = w.shape[0]
        g = g.contiguous()
        need_dx, need_dw, need_db, need_a, need_rz = ctx.needs_input_grad[:5]
        need_db = need_db and ctx.has_bias
        if need_dx and ctx.fork_state is not None:  # the fork node adds the skip's share to the convolution branch's
            ctx.fork_state.g, ctx.fork_state.w = g, w
            need_dx = False
        dx = torch.empty_like(x) if need_dx else None
        dw = torch.empty_like(w) if need_dw else None
        db = torch.empty(Fout, dtype=torch.float32, device=x.device) if need_db else None
        da = torch.empty_like(a) if need_a else None
        drz = torch.empty_like(rz) if need_rz else None
        with torch.cuda.device(x.device):
            if need_dx or need_dw or need_db:
                ws = _workspace(lib.dsw_linear_workspace_bytes(B, V, Fin, Fout), x.device)
                rc = lib.dsw_linear_bwd(
                    x.data_ptr(), x.stride(0), x.stride(1), g.data_ptr(), w.data_ptr(),
                    dx.data_ptr() if dx is not None else None, dw.data_ptr() if dw is not None else None,
                    db.data_ptr() if db is not None else None, B, V, Fin, Fout, ws.data_ptr(), ws.numel(), _stream_ptr(x.device))
                _lib.check(rc, "dsw_linear_bwd")
            if need_a or need_rz:
                ws2 = _workspace(lib.dsw_rezero_bwd_workspace_bytes(), x.device)
                rc = lib.dsw_rezero_bwd(g.data_ptr(), a.data_ptr(), rz.data_ptr(), da.data_ptr() if da is not None else None,
                                        drz.data_ptr() if drz is not None else None, ws2.data_ptr(), ws2.numel(), a.numel(),
                                        _stream_ptr(x.device))
                _lib.check(rc, "dsw_rezero_bwd")
        return dx, dw, db, da, drz, None, None


def linear_rezero(x, weight, bias, conv_out, rezero_weight, cat_slot=False, fork_state=None):
    return LinearRezeroFunction.apply(x, weight, bias, conv_out, rezero_weight, cat_slot, fork_state)


# --------------------------------------------------------------------------------------------
# Sparse remap                                                     reference layers.py:956-964
# --------------------------------------------------------------------------------------------


class RemapFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, plan: SparsePlan):
        _require_cuda_f32(x, "x")
        B, V, F = x.shape
        if V != plan.shape[1]:
            raise ValueError(f"x has {V} nodes but remap_matrix is {plan.shape}")
        _require_same_device(x, remap_matrix=plan)
        x = _channel_last(x)
        y = torch.empty((B, plan.shape[0], F), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().dsw_spmm_fwd(
                plan.handle, x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(), B, F, _stream_ptr(x.device)
            )
        _lib.check(rc, "dsw_spmm_fwd")
        ctx.plan = plan
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        plan = ctx.plan
        dy = _channel_last(dy)
        B, _, F = dy.shape
        dx = torch.empty((B, plan.shape[1], F), dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            rc = _lib.load().dsw_spmm_bwd(
                plan.handle, dy.data_ptr(), dy.stride(0), dy.stride(1), dx.data_ptr(), B, F, _stream_ptr(dy.device)
            )
        _lib.check(rc, "dsw_spmm_bwd")
        return dx, None


def remap(x, plan: SparsePlan):
    return RemapFunction.apply(x, plan)


# --------------------------------------------------------------------------------------------
# Skip concatenation without a concatenation pass        reference my_models_graph.py:505-538
# --------------------------------------------------------------------------------------------


def cat_slot_of(skip: torch.Tensor):
    """The ``[B, V, 2C]`` buffer whose second half ``skip`` is (``LinearRezeroFunction(cat_slot=True)``), or None."""
    if skip.dim() != 3 or not skip.is_cuda or skip.dtype != torch.float32:
        return None
    B, V, C = skip.shape
    if skip.stride() != (V * 2 * C, 2 * C, 1) or skip.storage_offset() != C:
        return None
    if skip.untyped_storage().nbytes() < B * V * 2 * C * 4:
        return None
    return skip.as_strided((B, V, 2 * C), (V * 2 * C, 2 * C, 1), 0)


class RemapCatFunction(torch.autograd.Function):
    """``torch.cat((remap(x), skip), dim=2)`` (``x = self.unpool(x, idx); torch.cat((x, x_enc), dim=2)``,
    ``my_models_graph.py:531-538``) where ``skip`` already occupies the second half of the result's buffer: the unpool
    writes its rows into the first half (``dsw_spmm_fwd_ex``, strided output) and the buffer itself is returned."""

    @staticmethod
    def forward(ctx, x, plan: SparsePlan, skip):
        _require_cuda_f32(x, "x")
        buf = cat_slot_of(skip)
        B, Vc, C = x.shape
        if buf is None or skip.shape[2] != C or skip.shape[0] != B or skip.shape[1] != plan.shape[0]:
            raise ValueError("skip is not the second half of a [B, V, 2C] concatenation buffer matching x")
        if Vc != plan.shape[1]:
            raise ValueError(f"x has {Vc} nodes but remap_matrix is {plan.shape}")
        _require_same_device(x, remap_matrix=plan, skip=skip)
        x = _channel_last(x)
        with torch.cuda.device(x.device):
            rc = _lib.load().dsw_spmm_fwd_ex(plan.handle, x.data_ptr(), x.stride(0), x.stride(1), None, 0, 0, buf.data_ptr(),
                                             buf.stride(0), buf.stride(1), B, C, _stream_ptr(x.device))
        _lib.check(rc, "dsw_spmm_fwd_ex")
        ctx.plan = plan
        ctx.C = C
        return buf

    @staticmethod
    @once_differentiable
    def backward(ctx, dcat):
        plan, C = ctx.plan, ctx.C
        dcat = _channel_last(dcat)
        B = dcat.shape[0]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((B, plan.shape[1], C), dtype=torch.float32, device=dcat.device)
            with torch.cuda.device(dcat.device):
                rc = _lib.load().dsw_spmm_bwd(plan.handle, dcat.data_ptr(), dcat.stride(0), dcat.stride(1), dx.data_ptr(), B, C,
                                              _stream_ptr(dcat.device))
            _lib.check(rc, "dsw_spmm_bwd")
        return dx, None, (dcat[:, :, C:] if ctx.needs_input_grad[2] else None)


class RemapForkFunction(torch.autograd.Function):
    """``(remap(x), x)``: the encoder output feeds the pool and the decoder's skip connection
    (``my_models_graph.py:505-511``).  One node owns both uses so that its backward adds the skip gradient inside the
    transposed product (``dsw_spmm_bwd_ex``: dx = M^T d_pooled + d_skip) instead of in a separate element-wise pass."""

    @staticmethod
    def forward(ctx, x, plan: SparsePlan):
        _require_cuda_f32(x, "x")
        B, V, F = x.shape
        if V != plan.shape[1]:
            raise ValueError(f"x has {V} nodes but remap_matrix is {plan.shape}")
        _require_same_device(x, remap_matrix=plan)
        xs = _channel_last(x)
        y = torch.empty((B, plan.shape[0], F), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().dsw_spmm_fwd(plan.handle, xs.data_ptr(), xs.stride(0), xs.stride(1), y.data_ptr(), B, F,
                                          _stream_ptr(x.device))
        _lib.check(rc, "dsw_spmm_fwd")
        ctx.plan = plan
        ctx.set_materialize_grads(False)
        return y, x.as_strided(x.shape, x.stride(), x.storage_offset())

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, dskip):
        plan = ctx.plan
        if dy is None:
            return dskip, None
        dy = _channel_last(dy)
        B, _, F = dy.shape
        gp, gsb, gsv = None, 0, 0
        if dskip is not None:
            if dskip.stride(2) != 1 or (dskip.stride(0) | dskip.stride(1) | dskip.storage_offset()) % 4:
                dskip = dskip.contiguous()
            gp, gsb, gsv = dskip.data_ptr(), dskip.stride(0), dskip.stride(1)
        dx = torch.empty((B, plan.shape[1], F), dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            rc = _lib.load().dsw_spmm_bwd_ex(plan.handle, dy.data_ptr(), dy.stride(0), dy.stride(1), gp, gsb, gsv, dx.data_ptr(),
                                             dx.stride(0), dx.stride(1), B, F, _stream_ptr(dy.device))
        _lib.check(rc, "dsw_spmm_bwd_ex")
        return dx, None


def remap_cat(x, plan: SparsePlan, skip):
    return RemapCatFunction.apply(x, plan, skip)


def remap_fork(x, plan: SparsePlan):
    return RemapForkFunction.apply(x, plan)


# --------------------------------------------------------------------------------------------
# Max-value pooling with indices                                 reference layers.py:1040-1103
# --------------------------------------------------------------------------------------------


class MaxValPoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, plan: SparsePlan):
        _require_cuda_f32(x, "x")
        B, V, F = x.shape
        assert V == plan.shape[1], "remap_matrix.shape[1] != x.shape[1]"  # layers.py:1048
        _require_same_device(x, remap_matrix=plan)
        Vc = plan.shape[0]
        x = _channel_last(x)
        y = torch.empty((B, Vc, F), dtype=torch.float32, device=x.device)
        index = torch.empty((2, F * B * Vc), dtype=torch.int64, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().dsw_maxval_pool_fwd(
                plan.handle, x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(),
                index[0].data_ptr(), index[1].data_ptr(), B, F, _stream_ptr(x.device),
            )
        _lib.check(rc, "dsw_maxval_pool_fwd")
        ctx.save_for_backward(index)
        ctx.dims = (B, V, Vc, F)
        ctx.mark_non_differentiable(index)
        return y, index

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _dindex):
        (index,) = ctx.saved_tensors
        B, V, Vc, F = ctx.dims
        dy = dy.contiguous()
        dx = torch.empty((B, V, F), dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            rc = _lib.load().dsw_maxval_pool_bwd(
                dy.data_ptr(), index[0].data_ptr(), dx.data_ptr(), B, V, Vc, F, _stream_ptr(dy.device)
            )
        _lib.check(rc, "dsw_maxval_pool_bwd")
        return dx, None


class ScatterUnpoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, index, n_fine: int):
        _require_cuda_f32(x, "x")
        B, Vc, F = x.shape
        if index.dtype != torch.int64 or index.shape != (2, F * B * Vc):
            raise ValueError("index must be the int64 [2, F*B*Vc] tensor returned by the max-value pool")
        x = x.contiguous()
        index = index.contiguous()
        out = torch.empty((B, n_fine, F), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().dsw_scatter_unpool_fwd(
                x.data_ptr(), index[0].data_ptr(), index[1].data_ptr(), out.data_ptr(), B, n_fine, Vc, F,
                _stream_ptr(x.device),
            )
        _lib.check(rc, "dsw_scatter_unpool_fwd")
        ctx.save_for_backward(index)
        ctx.dims = (B, n_fine, Vc, F)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (index,) = ctx.saved_tensors
        B, V, Vc, F = ctx.dims
        dout = dout.contiguous()
        dx = torch.empty((B, Vc, F), dtype=torch.float32, device=dout.device)
        with torch.cuda.device(dout.device):
            rc = _lib.load().dsw_scatter_unpool_bwd(
                dout.data_ptr(), index[0].data_ptr(), index[1].data_ptr(), dx.data_ptr(), B, V, Vc, F,
                _stream_ptr(dout.device),
            )
        _lib.check(rc, "dsw_scatter_unpool_bwd")
        return dx, None, None


# --------------------------------------------------------------------------------------------
# Nested-order pools                                               reference layers.py:784-941
# --------------------------------------------------------------------------------------------


class NestedMaxPoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel: int):
        _require_cuda_f32(x, "x")
        B, V, F = x.shape
        x = _channel_last(x)
        Vc = V // kernel
        y = torch.empty((B, Vc, F), dtype=torch.float32, device=x.device)
        idx = torch.empty((B, F, Vc), dtype=torch.int64, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().dsw_nested_maxpool_fwd(
                x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(), idx.data_ptr(), B, V, F, kernel,
                _stream_ptr(x.device),
            )
        _lib.check(rc, "dsw_nested_maxpool_fwd")
        ctx.save_for_backward(idx)
        ctx.dims = (B, V, F, kernel)
        ctx.mark_non_differentiable(idx)
        return y, idx

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _didx):
        (idx,) = ctx.saved_tensors
        B, V, F, kernel = ctx.dims
        dy = dy.contiguous()
        dx = torch.empty((B, V, F), dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            rc = _lib.load().dsw_nested_scatter(dy.data_ptr(), idx.data_ptr(), dx.data_ptr(), B, V, F, kernel,
                                               _stream_ptr(dy.device))
        _lib.check(rc, "dsw_nested_scatter")
        return dx, None


class NestedMaxUnpoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, kernel: int):
        _require_cuda_f32(x, "x")
        B, Vc, F = x.shape
        V = Vc * kernel
        x = x.contiguous()
        idx = idx.contiguous()
        out = torch.empty((B, V, F), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().dsw_nested_scatter(x.data_ptr(), idx.data_ptr(), out.data_ptr(), B, V, F, kernel,
                                               _stream_ptr(x.device))
        _lib.check(rc, "dsw_nested_scatter")
        ctx.save_for_backward(idx)
        ctx.dims = (B, V, F, kernel)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        B, V, F, kernel = ctx.dims
        dout = dout.contiguous()
        dx = torch.empty((B, V // kernel, F), dtype=torch.float32, device=dout.device)
        with torch.cuda.device(dout.device):
            rc = _lib.load().dsw_nested_gather(dout.data_ptr(), idx.data_ptr(), dx.data_ptr(), B, V, F, kernel,
                                              _stream_ptr(dout.device))
        _lib.check(rc, "dsw_nested_gather")
        return dx, None, None


class NestedAvgPoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel: int):
        _require_cuda_f32(x, "x")
        B, V, F = x.shape
        x = _channel_last(x)
        y = torch.empty((B, V // kernel, F), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().dsw_nested_avgpool_fwd(x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(), B, V, F,
                                                   kernel, _stream_ptr(x.device))
        _lib.check(rc, "dsw_nested_avgpool_fwd")
        ctx.dims = (B, V, F, kernel)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        B, V, F, kernel = ctx.dims
        dy = _channel_last(dy)
        dx = torch.empty((B, V, F), dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            rc = _lib.load().dsw_nested_repeat(dy.data_ptr(), dy.stride(0), dy.stride(1), dx.data_ptr(),
                                              1.0 / kernel, B, V, F, kernel, _stream_ptr(dy.device))
        _lib.check(rc, "dsw_nested_repeat")
        return dx, None


class NestedAvgUnpoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel: int):
        _require_cuda_f32(x, "x")
        B, Vc, F = x.shape
        x = _channel_last(x)
        V = Vc * kernel
        y = torch.empty((B, V, F), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().dsw_nested_repeat(x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(), 1.0, B, V, F,
                                              kernel, _stream_ptr(x.device))
        _lib.check(rc, "dsw_nested_repeat")
        ctx.dims = (B, V, F, kernel)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        B, V, F, kernel = ctx.dims
        dy = _channel_last(dy)
        dx = torch.empty((B, V // kernel, F), dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            rc = _lib.load().dsw_nested_sum(dy.data_ptr(), dy.stride(0), dy.stride(1), dx.data_ptr(), B, V, F,
                                           kernel, _stream_ptr(dy.device))
        _lib.check(rc, "dsw_nested_sum")
        return dx, None
