"""The autoregressive training step around ``model(X)``.

The reference hands its training loop to ``xforecasting.AutoregressiveTraining``
(``scripts_training/train_predict_state.py:392-436``): with ``ar_settings`` ``input_k = [-T .. -1]``, ``output_k = [0]``,
``forecast_cycle = 1`` and ``stack_most_recent_prediction`` (``modules/utils_config.py:82-86``) every batch runs
``ar_iterations + 1`` forward passes; iteration ``i`` stacks ``X = cat(dynamic history, boundary conditions, static)`` —
the newest history slots being the model's own previous predictions —, calls ``Y = model(X)``, and adds
``w_i * criterion(reshape(Y), reshape(Y_obs_i))`` (``modules/loss.py:30-53, 118-160``) to the loss that is back-propagated
once.  ``xforecasting`` is not installable here; this module restates that loop on the CUDA library:

* :func:`ar_stack` — one kernel (``dsw_ar_stack_*``) writes ``X`` from per-time-slot *pointers* (observed states or earlier
  predictions), instead of a shifted copy of the history, an expand of the static fields and a three-way ``cat``;
* :class:`ARRollout` — the loop; :meth:`ARRollout.capture` records forward + backward of the whole rollout into ONE CUDA
  graph (at 8 samples per GPU the ~2 000 kernel launches of a 3-iteration rollout otherwise bound the step on the host).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from .functional import _require_cuda_f32, _require_same_device, _stream_ptr

MAX_SLOTS = 8


class _ArSlots(C.Structure):
    _fields_ = [
        ("dyn", C.c_void_p * MAX_SLOTS), ("dyn_sB", C.c_int64 * MAX_SLOTS), ("dyn_sV", C.c_int64 * MAX_SLOTS),
        ("bc", C.c_void_p * MAX_SLOTS), ("bc_sB", C.c_int64 * MAX_SLOTS), ("bc_sV", C.c_int64 * MAX_SLOTS),
        ("stat", C.c_void_p), ("ddyn", C.c_void_p * MAX_SLOTS),
    ]


def _unit_feature_stride(t: torch.Tensor) -> torch.Tensor:
    return t if (t.stride(2) == 1 or t.shape[2] == 1) and t.stride(2) in (0, 1) or t.shape[2] == 1 else t.contiguous()


class ArStackFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, static, T: int, *slots):
        dyn, bc = slots[:T], slots[T:]
        B, V, Fd = dyn[0].shape
        Fb = bc[0].shape[2] if bc else 0
        Fs = static.shape[1] if static is not None else 0
        if T > MAX_SLOTS:
            raise ValueError(f"at most {MAX_SLOTS} input time slots")
        s = _ArSlots()
        keep = []
        for t in range(T):
            d = dyn[t]
            _require_cuda_f32(d, f"dynamic slot {t}")
            _require_same_device(dyn[0], **{f"dynamic slot {t}": d})
            if tuple(d.shape) != (B, V, Fd):
                raise ValueError(f"dynamic slot {t} has shape {tuple(d.shape)}, expected {(B, V, Fd)}")
            if d.stride(2) != 1 and Fd > 1:
                d = d.contiguous()
            keep.append(d)
            s.dyn[t], s.dyn_sB[t], s.dyn_sV[t] = d.data_ptr(), d.stride(0), d.stride(1)
            if Fb:
                c = bc[t]
                _require_cuda_f32(c, f"boundary-condition slot {t}")
                _require_same_device(dyn[0], **{f"boundary-condition slot {t}": c})
                if tuple(c.shape) != (B, V, Fb):
                    raise ValueError(f"boundary-condition slot {t} has shape {tuple(c.shape)}, expected {(B, V, Fb)}")
                if c.stride(2) != 1 and Fb > 1:
                    c = c.contiguous()
                keep.append(c)
                s.bc[t], s.bc_sB[t], s.bc_sV[t] = c.data_ptr(), c.stride(0), c.stride(1)
        if Fs:
            _require_cuda_f32(static, "static")
            _require_same_device(dyn[0], static=static)
            if static.shape[0] != V:
                raise ValueError(f"static has {static.shape[0]} nodes, expected {V}")
            st = static.contiguous()
            keep.append(st)
            s.stat = st.data_ptr()
        X = torch.empty((B, T, V, Fd + Fb + Fs), dtype=torch.float32, device=dyn[0].device)
        with torch.cuda.device(X.device):
            rc = _lib.load().dsw_ar_stack_fwd(C.byref(s), X.data_ptr(), B, T, V, Fd, Fb, Fs, _stream_ptr(X.device))
        _lib.check(rc, "dsw_ar_stack_fwd")
        ctx.dims = (B, T, V, Fd, Fb, Fs)
        return X

    @staticmethod
    @once_differentiable
    def backward(ctx, dX):
        B, T, V, Fd, Fb, Fs = ctx.dims
        dX = dX.contiguous()
        s = _ArSlots()
        grads = []
        for t in range(T):
            if ctx.needs_input_grad[2 + t]:
                g = torch.empty((B, V, Fd), dtype=torch.float32, device=dX.device)
                s.ddyn[t] = g.data_ptr()
                grads.append(g)
            else:
                grads.append(None)
        if any(g is not None for g in grads):
            with torch.cuda.device(dX.device):
                rc = _lib.load().dsw_ar_stack_bwd(C.byref(s), dX.data_ptr(), B, T, V, Fd, Fb, Fs, _stream_ptr(dX.device))
            _lib.check(rc, "dsw_ar_stack_bwd")
        n_bc = T if Fb else 0
        return (None, None, *grads, *([None] * n_bc))


def ar_stack(dyn_slots: Sequence[torch.Tensor], bc_slots: Optional[Sequence[torch.Tensor]] = None,
             static: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``X[b, t, v, :] = [dyn_slots[t][b, v, :], bc_slots[t][b, v, :], static[v, :]]`` — the model input of one
    autoregressive iteration, ``[sample, time, node, feature]``.  Slots are ``[B, V, F]`` tensors (views welcome: any batch /
    node strides); gradients flow to the dynamic slots only (the predictions of earlier iterations)."""
    T = len(dyn_slots)
    bc_slots = list(bc_slots) if bc_slots is not None else []
    if bc_slots and len(bc_slots) != T:
        raise ValueError("need one boundary-condition slot per input time slot")
    return ArStackFunction.apply(static, T, *dyn_slots, *bc_slots)


class ARRollout:
    """``ar_iterations + 1`` forward passes with the model's predictions fed back, one backward pass.

    ``history``  [B, T, V, Fd]   observed dynamic states at times -T .. -1
    ``bc``       [B, T + n, V, Fb] boundary conditions at times -T .. n - 1 (or None), ``n = ar_iterations + 1``
    ``static``   [V, Fs] (or None)
    ``targets``  [B, n, V, Fd]   observed dynamic states at times 0 .. n - 1
    The model maps ``[B, T, V, Fd + Fb + Fs]`` to ``[B, 1, V, Fd]`` (``UNetSpherical`` with ``output_n_time = 1``).
    """

    def __init__(self, model: torch.nn.Module, criterion, ar_iterations: int, ar_weights: Optional[Sequence[float]] = None):
        self.model, self.criterion = model, criterion
        self.n = int(ar_iterations) + 1
        self.weights = [float(w) for w in (ar_weights if ar_weights is not None else [1.0 / self.n] * self.n)]
        if len(self.weights) != self.n:
            raise ValueError("need one loss weight per autoregressive iteration")
        self._graph = None

    def loss(self, history, bc, static, targets):
        T = history.shape[1]
        states = [history[:, t] for t in range(T)]  # views; predictions are appended
        total = None
        for i in range(self.n):
            X = ar_stack(states[i:i + T], [bc[:, i + t] for t in range(T)] if bc is not None else None, static)
            Y = self.model(X)                      # [B, 1, V, Fd]
            pred = Y[:, 0]
            li = self.criterion(pred, targets[:, i])   # reshape_tensors_4_loss: data points = samples (one output time)
            total = li * self.weights[i] if total is None else total + li * self.weights[i]
            states.append(pred)
        return total

    def step(self, history, bc, static, targets, zero_grad=None):
        """Forward + backward of one batch; returns the (detached) loss.  ``zero_grad`` is called between the forward and
        the backward pass (``optimizer.zero_grad`` / ``FlatGradBucket.zero_``)."""
        total = self.loss(history, bc, static, targets)
        if zero_grad is not None:
            zero_grad()
        total.backward()
        return total.detach()

    # ---- CUDA graph: the whole rollout (forward x n, losses, backward) as one replay ----
    def capture(self, history, bc, static, targets, zero_grad=None, warmup: int = 3):
        """Record :meth:`step` into a CUDA graph on static copies of the arguments.  Afterwards :meth:`replay` copies new
        batches into those buffers and replays; parameter gradients land in the same ``p.grad`` tensors every time.
        (PyTorch's rule for whole-step capture applies: no loss tensor of an earlier, eager step may still be alive —
        it would keep AccumulateGrad nodes of the default stream in the graph.  :meth:`step` returns a detached loss.)"""
        import gc

        gc.collect()
        dev = history.device
        self._static = [t.clone() if t is not None else None for t in (history, bc, static, targets)]
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(*self._static, zero_grad=zero_grad)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._loss = self.step(*self._static, zero_grad=zero_grad)
        return self

    def replay(self, history, bc, static, targets):
        if self._graph is None:
            raise RuntimeError("call capture() first")
        for dst, src in zip(self._static, (history, bc, static, targets)):
            if dst is not None and src is not None and dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._loss
