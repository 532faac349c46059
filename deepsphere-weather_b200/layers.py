"""Drop-in layers with the reference's L2 interface (``modules/layers.py`` of
deepsphere/deepsphere-weather), computing through ``libdsw.so``.

Same class names, constructor signatures, attributes, buffers and state-dict keys as the
reference, so ``my_models_graph.py`` / ``models.py`` run on them unchanged (INTEGRATION.md shows
the one-line rebinding).  Differences by design:

* there is no CPU path — CUDA tensors only;
* pool / unpool outputs are contiguous ``[B, V', F]`` (the reference returns a strided view of a
  ``[V', F, B]`` buffer, ``layers.py:963``); values are identical;
* sparse operators are turned into device plans lazily on the first forward after ``.to(device)``
  and are *not* part of the state dict.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from scipy import sparse

from . import functional as F_
from .graphs import prepare_torch_laplacian, scipy_to_torch_coo  # noqa: F401  (re-export)

_RELU_LIKE = {
    "relu", "celu", "selu", "prelu", "hardswish", "mish", "silu", "gelu", "softplus", "softmax",
    "logsigmoid", "relu6", "rrlu", "leaky_relu", "elu",
}
_LINEAR_LIKE = {"linear", "hardshrink ", "sigmoid", "hardsigmoid", "tanh", "hardtanh", "softsign"}


def convert_to_torch_sparse(mat) -> torch.Tensor:
    """scipy sparse -> coalesced torch COO (reference ``layers.py:584-594``).  A torch sparse tensor — e.g. an operator
    built on the device by ``graphs_device`` — is taken as it is."""
    if torch.is_tensor(mat):
        if not mat.is_sparse:
            raise TypeError("expected a scipy sparse matrix or a torch sparse COO tensor")
        return mat.coalesce().to(torch.get_default_dtype())
    return scipy_to_torch_coo(mat, torch.get_default_dtype())


# ------------------------------------------------------------------------------------------
# Chebyshev graph convolution
# ------------------------------------------------------------------------------------------


def conv_cheb(laplacian: torch.Tensor, inputs: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """Functional form with the reference's signature (``layers.py:113``)."""
    return _cheb_conv_padded(inputs, weight, None, F_.plan_for(laplacian))


def _pad4(n: int) -> int:
    return (-n) % 4


def _cheb_conv_padded(inputs, weight, bias, plan, act=0, input_is_relu=False, premasked=False):
    """Channel counts that are not multiples of 4 (21 input features, 2 outputs) would force the
    unaligned scalar kernels (no 16-byte rows, no TMA).  Zero-padding the channel dimension of the
    activations and of ``weight`` / ``bias`` is exact — padded inputs meet zero weights, padded outputs
    are sliced away — and lets every layer use the vectorised / tensor-map paths; autograd slices the
    gradients back."""
    pin, pout = _pad4(inputs.shape[2]), _pad4(weight.shape[2])
    if inputs.shape[2] != weight.shape[0] or (pin == 0 and pout == 0):
        # (a shape mismatch raises the reference's error there)
        return F_.cheb_conv(inputs, weight, bias, plan, act, input_is_relu, premasked)
    if input_is_relu or (premasked and pout):
        raise ValueError("the ReLU-mask delegation needs channel counts that are multiples of 4 (see ConvCheb.relu_chain_ok)")
    x = torch.nn.functional.pad(inputs, (0, pin)) if pin else inputs
    w = torch.nn.functional.pad(weight, (0, pout, 0, 0, 0, pin)) if (pin or pout) else weight
    b = torch.nn.functional.pad(bias, (0, pout)) if (bias is not None and pout) else bias
    y = F_.cheb_conv(x, w, b, plan, act, False, premasked)  # (a padded INPUT does not touch the output the next layer masks by)
    return y[..., : weight.shape[2]] if pout else y


class ConvCheb(torch.nn.Module):
    """Graph convolution by Chebyshev polynomials of the Laplacian (reference ``layers.py:183-376``).

    ``forward(inputs[B, V, Fin]) -> [B, V, Fout]``; parameters ``weight[Fin, K, Fout]``,
    ``bias[Fout]`` (or ``None``); buffer ``laplacian`` (sparse COO, persistent — checkpoints of the
    reference load with ``strict=True``).  Recurrence, channel mix and bias run as one library call.
    """

    def __init__(self, in_channels, out_channels, kernel_size, laplacian, bias=True, conv=conv_cheb, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self._conv = conv
        self.register_buffer("laplacian", laplacian)
        self.weight = torch.nn.Parameter(torch.empty(in_channels, kernel_size, out_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self, activation="relu", fan="in", distribution="normal"):
        """He / Glorot / LeCun initialisation, same options as the reference (``layers.py:253-343``)."""
        k = self.kernel_size
        if fan == "in":
            fan_v = self.in_channels * k
        elif fan == "out":
            fan_v = self.out_channels * k
        elif fan == "avg":
            fan_v = (self.in_channels + self.out_channels) / 2 * k
        else:
            raise ValueError("unknown fan")
        if activation in _RELU_LIKE:
            gain = 2
        elif activation in _LINEAR_LIKE:
            gain = 1
        else:
            raise ValueError("Unknown activation")
        with torch.no_grad():
            if distribution == "normal":
                self.weight.normal_(0, math.sqrt(gain / fan_v))
            elif distribution == "uniform":
                lim = math.sqrt(3 * gain / fan_v)
                self.weight.uniform_(-lim, lim)
            else:
                raise ValueError("Unknown distribution")
            if self.bias is not None:
                self.bias.fill_(0)

    def set_parameters(self, weight, bias=None):
        self.weight = torch.nn.Parameter(torch.as_tensor(weight))
        if bias is not None:
            self.bias = torch.nn.Parameter(torch.as_tensor(bias))

    def extra_repr(self):
        return "{} -> {}, kernel_size={}, bias={}".format(
            self.in_channels, self.out_channels, self.kernel_size, self.bias is not None
        )

    # activations ``forward(..., activation=)`` can fuse into the convolution's last kernel
    fused_activations = ("relu",)

    def relu_chain_ok(self, consumer: bool = True) -> bool:
        """True when this layer can take part in the ReLU-mask delegation of ``forward(..., input_is_relu= / premasked=)``:
        library convolution whose output (producer, ``premasked``) — and, for the masking consumer (``input_is_relu``),
        input as well — needs no channel padding."""
        return (self._conv is conv_cheb and _pad4(self.out_channels) == 0
                and (not consumer or _pad4(self.in_channels) == 0))

    def forward(self, inputs, activation=None, input_is_relu=False, premasked=False):
        """``forward(inputs)`` is the reference's signature; ``activation="relu"`` (an extension used by
        ``models.ConvBlock``) applies the ReLU inside the convolution's last kernel instead of in a
        separate element-wise pass.

        ``input_is_relu`` / ``premasked`` (extensions used by ``models.ResBlock``) move the ReLU's backward from a
        separate element-wise pass into the NEXT layer's input-gradient kernel: a layer called with
        ``activation="relu", premasked=True`` promises that its output feeds exactly one layer, which is called with
        ``input_is_relu=True`` and returns the gradient already multiplied by ``[x > 0]``."""
        if activation not in (None,) + self.fused_activations:
            raise ValueError(f"activation {activation!r} cannot be fused; apply it to the output instead")
        if self._conv is not conv_cheb:  # user-supplied convolution: honour the reference contract
            if input_is_relu or premasked:
                raise ValueError("the ReLU-mask delegation needs the library convolution")
            out = self._conv(self.laplacian, inputs, self.weight)
            if self.bias is not None:
                out += self.bias
            return torch.relu(out) if activation == "relu" else out
        if premasked and activation != "relu":
            raise ValueError("premasked=True only makes sense with activation='relu'")
        return _cheb_conv_padded(inputs, self.weight, self.bias, F_.plan_for(self.laplacian), 1 if activation == "relu" else 0,
                                 input_is_relu, premasked)


class NodeLinear(torch.nn.Linear):
    """``torch.nn.Linear`` (same parameters, initialisation and state-dict keys) whose forward /
    backward on ``[B, V, F]`` node features run on the tcgen05 kernels — the ResBlock skip connection
    of the reference (``my_models_graph.py:196-201``).  Other input ranks fall through to torch."""

    def forward(self, x):
        if x.dim() == 3 and x.is_cuda and x.dtype == torch.float32:
            pin, pout = _pad4(self.in_features), _pad4(self.out_features)
            if x.shape[2] != self.in_features or (pin == 0 and pout == 0):
                return F_.NodeLinearFunction.apply(x, self.weight, self.bias)
            # zero-pad unaligned channel counts (exact; see _cheb_conv_padded)
            xp = torch.nn.functional.pad(x, (0, pin)) if pin else x
            w = torch.nn.functional.pad(self.weight, (0, pin, 0, pout))
            b = torch.nn.functional.pad(self.bias, (0, pout)) if (self.bias is not None and pout) else self.bias
            y = F_.NodeLinearFunction.apply(xp, w, b)
            return y[..., : self.out_features] if pout else y
        return super().forward(x)

    def fusable_rezero(self, x) -> bool:
        """True when ``forward_rezero`` runs as the single fused launch on ``x`` itself (16-byte aligned channel counts)."""
        return (x.dim() == 3 and x.is_cuda and x.dtype == torch.float32 and _pad4(self.in_features) == 0
                and _pad4(self.out_features) == 0)

    def forward_rezero(self, x, conv_out, rezero_weight, cat_slot=False, fork_state=None):
        """``self(x) + rezero_weight * conv_out`` — the ResBlock tail (``my_models_graph.py:211-215``) — in one
        launch when the output channel count is 16-byte aligned (an unaligned input count is zero-padded first, which is
        exact), else the two-kernel composition.  ``cat_slot`` / ``fork_state``: see ``functional.LinearRezeroFunction``
        (only honoured by the fused launch; a fork needs aligned inputs)."""
        if (x.dim() == 3 and x.is_cuda and x.dtype == torch.float32 and conv_out.dtype == torch.float32
                and _pad4(self.out_features) == 0 and x.shape[2] == self.in_features):
            pin = _pad4(self.in_features)
            if pin == 0:
                return F_.linear_rezero(x, self.weight, self.bias, conv_out, rezero_weight, cat_slot, fork_state)
            if fork_state is None:
                pad = torch.nn.functional.pad
                return F_.linear_rezero(pad(x, (0, pin)), pad(self.weight, (0, pin)), self.bias, conv_out, rezero_weight,
                                        cat_slot, None)
        return F_.rezero_residual(conv_out, self.forward(x), rezero_weight)


# ------------------------------------------------------------------------------------------
# Nested-order (HEALPix) pools                                    reference layers.py:784-941
# ------------------------------------------------------------------------------------------


class HealpixMaxPool(torch.nn.Module):
    def __init__(self, kernel_size, return_indices=True, *args, **kwargs):
        super().__init__()
        self.kernel_size = kernel_size
        self.return_indices = return_indices

    def extra_repr(self):
        return f"kernel_size={self.kernel_size}"

    def forward(self, x):
        y, idx = F_.NestedMaxPoolFunction.apply(x, self.kernel_size)
        return (y, idx) if self.return_indices else y


class HealpixMaxUnpool(torch.nn.Module):
    def __init__(self, kernel_size, *args, **kwargs):
        super().__init__()
        self.kernel_size = kernel_size

    def extra_repr(self):
        return f"kernel_size={self.kernel_size}"

    def forward(self, x, indices, **kwargs):
        return F_.NestedMaxUnpoolFunction.apply(x, indices, self.kernel_size)


class HealpixAvgPool(torch.nn.Module):
    def __init__(self, kernel_size, *args, **kwargs):
        super().__init__()
        self.kernel_size = kernel_size

    def extra_repr(self):
        return f"kernel_size={self.kernel_size}"

    def forward(self, x):
        return F_.NestedAvgPoolFunction.apply(x, self.kernel_size), None


class HealpixAvgUnpool(torch.nn.Module):
    def __init__(self, kernel_size, *args, **kwargs):
        super().__init__()
        self.kernel_size = kernel_size

    def extra_repr(self):
        return f"kernel_size={self.kernel_size}"

    def forward(self, x, *args):
        return F_.NestedAvgUnpoolFunction.apply(x, self.kernel_size)


# ------------------------------------------------------------------------------------------
# Remap-matrix pools                                             reference layers.py:948-1103
# ------------------------------------------------------------------------------------------


class RemapBlock(torch.nn.Module):
    """``out[b, v', f] = sum_v M[v', v] x[b, v, f]`` with ``M`` the ``remap_matrix`` buffer."""

    def __init__(self, remap_matrix):
        super().__init__()
        self.register_buffer("remap_matrix", self.process_remap_matrix(remap_matrix))

    def process_remap_matrix(self, mat):
        return convert_to_torch_sparse(mat)

    def forward(self, x, *args, **kwargs):
        return F_.remap(x, F_.plan_for(self.remap_matrix))


def pool_fork(pool, x):
    """``(pooled, indices, skip)`` with ``skip`` the encoder output to hand to the decoder (``my_models_graph.py:505-511``:
    ``x_enc = conv(x); x, idx = pool(x_enc)``).  For the sparse-remap pools one autograd node owns both uses, so the
    backward adds the skip gradient inside the transposed product; other pools get the plain composition."""
    if type(pool) in (GeneralAvgPool, GeneralMaxAreaPool) and x.is_cuda and x.dim() == 3 and x.dtype == torch.float32:
        pooled, skip = F_.remap_fork(x, F_.plan_for(pool.remap_matrix))
        return pooled, None, skip
    pooled, idx = pool(x)
    return pooled, idx, x


def unpool_cat(unpool, x, idx, skip):
    """``torch.cat((unpool(x, idx), skip), dim=2)`` (``my_models_graph.py:531-538``).  When ``skip`` was written into the
    second half of a ``[B, V, 2C]`` buffer by its ResBlock (``cat_slot=True``) and the unpool is a sparse remap, the
    unpool writes the first half in place and no concatenation pass runs."""
    if (type(unpool) in (GeneralAvgUnpool, GeneralMaxAreaUnpool) and x.is_cuda and x.dim() == 3 and x.dtype == torch.float32
            and x.shape[2] == skip.shape[2] and F_.cat_slot_of(skip) is not None):
        return F_.remap_cat(x, F_.plan_for(unpool.remap_matrix), skip)
    return torch.cat((unpool(x, idx), skip), dim=2)


class GeneralAvgPool(RemapBlock):
    def forward(self, x, *args, **kwargs):
        return super().forward(x), None


class GeneralAvgUnpool(RemapBlock):
    def forward(self, x, *args, **kwargs):
        return super().forward(x)


def _one_hot_coo(rows, cols, shape) -> torch.Tensor:
    idx = torch.from_numpy(np.stack([np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)]))
    return torch.sparse_coo_tensor(
        idx, torch.ones(idx.shape[1]), tuple(shape), dtype=torch.get_default_dtype(), check_invariants=False
    ).coalesce()


class GeneralMaxAreaPool(RemapBlock):
    """Each coarse node copies the fine node with which it shares the largest area
    (reference ``layers.py:991-1015``)."""

    def forward(self, x, *args, **kwargs):
        return super().forward(x), None

    def process_remap_matrix(self, mat):
        m = sparse.csr_matrix(mat)
        winners = np.asarray(m.argmax(axis=1)).ravel()
        return _one_hot_coo(np.arange(m.shape[0]), winners, m.shape)


class GeneralMaxAreaUnpool(RemapBlock):
    """Takes ``pool.T`` (fine x coarse); each coarse node writes to the single fine node holding
    its column maximum (reference ``layers.py:1018-1036``)."""

    def process_remap_matrix(self, mat):
        m = sparse.csc_matrix(mat)
        winners = np.asarray(m.argmax(axis=0)).ravel()
        return _one_hot_coo(winners, np.arange(m.shape[1]), m.shape)


class GeneralMaxValPool(RemapBlock):
    """Per coarse node and channel, keep the fine value whose *weighted* value is largest; also
    returns the reference's ``nnz_ind`` index tensor (``layers.py:1040-1083``)."""

    def forward(self, x, *args, **kwargs):
        return F_.MaxValPoolFunction.apply(x, F_.plan_for(self.remap_matrix))


class GeneralMaxValUnpool(RemapBlock):
    """Scatter pooled values back to the fine nodes they came from (``layers.py:1086-1103``)."""

    def forward(self, x, index, *args, **kwargs):
        return F_.ScatterUnpoolFunction.apply(x, index, self.remap_matrix.shape[0])


# ------------------------------------------------------------------------------------------
# Factories                                                       reference layers.py:1139-1242
# ------------------------------------------------------------------------------------------

from .layers_equiangular import (EQUIANGULAR_POOL, Conv2dEquiangular, EquiangularAvgPool, EquiangularAvgUnpool,  # noqa: E402,F401
                                 EquiangularMaxPool, EquiangularMaxUnpool)

HEALPIX_POOL = {"max": (HealpixMaxPool, HealpixMaxUnpool), "avg": (HealpixAvgPool, HealpixAvgUnpool)}
EQUIANGULAR_POOl = EQUIANGULAR_POOL  # (the reference's spelling, layers.py:1144)
ALL_POOL = {"healpix": HEALPIX_POOL, "equiangular": EQUIANGULAR_POOL}


class PoolUnpoolBlock(torch.nn.Module):
    @staticmethod
    def getPoolUnpoolLayer(sampling: str, pool_method: str, **kwargs):
        sampling, pool_method = sampling.lower(), pool_method.lower()
        assert sampling in ("healpix", "equiangular")
        assert pool_method in ("max", "avg")
        pool, unpool = ALL_POOL[sampling][pool_method]
        return pool(**kwargs), unpool(**kwargs)

    @staticmethod
    def getGeneralPoolUnpoolLayer(src_graph=None, dst_graph=None, pool_method: str = "interp", matrices=None):
        """``matrices = (pool_mat, unpool_mat)`` as scipy sparse: the reference computes them with
        xsphere/CDO (``layers.py:576-581``), which is an input to — not part of — the hot path."""
        if matrices is None:
            raise NotImplementedError("conservative-remap weights come from xsphere/CDO; pass matrices=(pool, unpool)")
        pool_mat, unpool_mat = matrices
        if pool_method == "interp":
            return GeneralAvgPool(pool_mat), GeneralAvgUnpool(unpool_mat)
        if pool_method == "maxarea":
            return GeneralMaxAreaPool(pool_mat), GeneralMaxAreaUnpool(sparse.coo_matrix(pool_mat).T)
        if pool_method == "maxval":
            return GeneralMaxValPool(pool_mat), GeneralMaxValUnpool(unpool_mat)
        if pool_method == "learn":
            raise NotImplementedError()
        raise ValueError(f"{pool_method} is not supoorted.")


def get_conv_fun(conv_type):
    """``{"image": Conv2dEquiangular, "graph": ConvCheb}`` (reference ``layers.py:1198-1201``)."""
    return {"image": Conv2dEquiangular, "graph": ConvCheb}[conv_type]


class GeneralConvBlock(torch.nn.Module):
    @staticmethod
    def getConvLayer(in_channels: int, out_channels: int, kernel_size: int, conv_type: str = "graph", **kwargs):
        """The layer for ``conv_type`` with the reference's keyword routing (``layers.py:1212-1242``)."""
        conv_type = conv_type.lower()
        if conv_type == "graph":
            assert "laplacian" in kwargs
            kwargs.pop("lonlat_ratio", None)
            kwargs.pop("periodic_padding", None)
            return get_conv_fun(conv_type)(in_channels, out_channels, kernel_size, **kwargs)
        if conv_type == "image":
            assert "lonlat_ratio" in kwargs
            kwargs.pop("laplacian", None)
            return get_conv_fun(conv_type)(in_channels, out_channels, kernel_size, **kwargs)
        raise ValueError("{} conv_type is not supported. Choose either 'graph' or 'image'".format(conv_type))
