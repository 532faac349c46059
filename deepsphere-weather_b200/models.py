"""Callers of the hot path: ``ConvBlock`` / ``ResBlock`` / ``UNetSpherical`` with the reference's
constructor signatures, module names and state-dict keys (reference
``modules/my_models_graph.py:26-564``, ``modules/models.py:16-120``).

The reference builds its graphs with pygsp and its interpolation-pool weights with CDO; neither
exists here, so graphs come from ``graphs.py`` and general pools from exact
nested-pixel matrices or caller-supplied ones (``pool_matrices=``).  Everything else — channel
plan, ReZero residuals, skip concatenation, tensor reshapes — is the reference architecture so
the whole-model parity test can compare outputs and gradients with the unmodified reference.

``backend`` lets tests / the CPU-baseline leg build the *same* architecture on other layer
classes (e.g. the oracle's); the product default is ``deepsphere_weather_b200.layers``.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Dict, Optional, Sequence

import numpy as np
import torch
from torch.nn import functional as F

from . import graphs as G


# A/B switch (DSW_LINEAR_REZERO=0): two-kernel ResBlock tail instead of the fused Linear + ReZero launch
_LINEAR_REZERO = os.environ.get("DSW_LINEAR_REZERO", "1") != "0"
# A/B switch (DSW_FUSED_SKIPS=0): torch.cat for the skip concatenation and autograd's own gradient accumulation adds
# instead of the cat-slot buffers / fork nodes
_FUSED_SKIPS = os.environ.get("DSW_FUSED_SKIPS", "1") != "0"
_FORK = os.environ.get("DSW_FORK", "1") != "0"  # A/B: fork nodes only


def _default_backend():
    from . import functional as F_
    from . import layers as L

    return SimpleNamespace(
        ConvCheb=L.ConvCheb,
        Linear=L.NodeLinear,
        healpix_pools={"max": (L.HealpixMaxPool, L.HealpixMaxUnpool), "avg": (L.HealpixAvgPool, L.HealpixAvgUnpool)},
        general_pools=L.PoolUnpoolBlock.getGeneralPoolUnpoolLayer,
        rezero_residual=F_.rezero_residual,
        fork=F_.fork, pool_fork=L.pool_fork, unpool_cat=L.unpool_cat,
    )


class ConvBlock(torch.nn.Module):
    """conv -> [BN] -> activation -> [BN]   (reference ``my_models_graph.py:26-118``)."""

    def __init__(self, in_channels, out_channels, laplacian, kernel_size=3, conv_type="graph", bias=True,
                 batch_norm=False, batch_norm_before_activation=False, activation=True, activation_fun="relu",
                 periodic_padding=True, lonlat_ratio=2, backend=None):
        super().__init__()
        backend = backend or _default_backend()
        if batch_norm:
            bias = False
        if conv_type == "graph":
            self.conv = backend.ConvCheb(in_channels, out_channels, kernel_size, laplacian=laplacian, bias=bias)
        elif conv_type == "image":  # the reference's dense lat x lon convolution (cuDNN)
            from .layers_equiangular import Conv2dEquiangular

            self.conv = Conv2dEquiangular(in_channels, out_channels, kernel_size, lonlat_ratio=lonlat_ratio,
                                          periodic_padding=periodic_padding, bias=bias)
        else:
            raise ValueError("{} conv_type is not supported. Choose either 'graph' or 'image'".format(conv_type))
        if batch_norm:
            self.bn = torch.nn.BatchNorm1d(out_channels)
        self.bn_before_act = batch_norm_before_activation
        self.norm = batch_norm
        self.act = activation
        self.act_fun = getattr(F, activation_fun)
        # ReLU directly after the convolution: let the convolution apply it in its last kernel
        self._fused_act = (activation_fun if activation and activation_fun in getattr(self.conv, "fused_activations", ())
                           and not (batch_norm and batch_norm_before_activation) else None)

    def _bn(self, x):
        return self.bn(x.permute(0, 2, 1)).permute(0, 2, 1)

    def relu_chain_ok(self, consumer: bool = True) -> bool:
        """This block is a library convolution (+ fused ReLU) and nothing else, on aligned channel counts: it can take
        part in the ReLU-mask delegation (``layers.ConvCheb.forward``)."""
        ok = getattr(self.conv, "relu_chain_ok", None)
        return bool(ok and ok(consumer) and not self.norm and (self._fused_act == "relu" or not self.act))

    def forward(self, x, input_is_relu=False, premasked=False):
        if input_is_relu or premasked:  # (only requested by ResBlock after relu_chain_ok())
            return self.conv(x, activation=self._fused_act, input_is_relu=input_is_relu, premasked=premasked)
        if self._fused_act is not None:
            x = self.conv(x, activation=self._fused_act)
            return self._bn(x) if self.norm else x
        x = self.conv(x)
        if self.norm and self.bn_before_act:
            x = self._bn(x)
        if self.act:
            x = self.act_fun(x)
        if self.norm and not self.bn_before_act:
            x = self._bn(x)
        return x


class ResBlock(torch.nn.Module):
    """n ConvBlocks (last one without activation), ReZero scalar, linear/identity skip
    (reference ``my_models_graph.py:121-216``)."""

    def __init__(self, in_channels, out_channels, laplacian, convblock_kwargs, **kwargs):
        super().__init__()
        self.rezero = True
        if not isinstance(out_channels, (int, tuple, list)):
            raise TypeError("'output_channels' must be int or list/tuple of int.")
        widths = list(out_channels) if isinstance(out_channels, (tuple, list)) else [out_channels]
        self.conv_names_list = []
        width_in = in_channels
        for i, width_out in enumerate(widths, start=1):
            kw = dict(convblock_kwargs)
            if i == len(widths):
                kw["activation"] = False
            name = f"convblock{i}"
            setattr(self, name, ConvBlock(width_in, width_out, laplacian=laplacian, **kw))
            self.conv_names_list.append(name)
            width_in = width_out
        if in_channels == widths[-1]:
            self.res_connection = torch.nn.Identity()
        else:
            backend = convblock_kwargs.get("backend") or _default_backend()
            self.res_connection = getattr(backend, "Linear", torch.nn.Linear)(in_channels, widths[-1])
        if self.rezero:
            self.rezero_weight = torch.nn.Parameter(torch.zeros(1), requires_grad=True)
        # fused `out * rezero_weight + skip` of the backend, if it has one
        backend = convblock_kwargs.get("backend") or _default_backend()
        self._fused_tail = getattr(backend, "rezero_residual", None)
        self._fork = getattr(backend, "fork", None) if (_FUSED_SKIPS and _FORK) else None
        if convblock_kwargs.get("batch_norm", False):
            last = getattr(self, self.conv_names_list[-1])
            torch.nn.init.constant_(last.bn.weight, 0)
            torch.nn.init.constant_(last.bn.bias, 0)

    def _run_convs(self, out):
        """The block's ConvBlocks in sequence.  Where block i ends in a fused ReLU and block i + 1 is a plain library
        convolution, the ReLU's backward is delegated to block i + 1's input-gradient kernel (no threshold pass)."""
        blocks = [getattr(self, name) for name in self.conv_names_list]
        relu_in = False
        for i, blk in enumerate(blocks):
            delegate = (_FUSED_SKIPS and i + 1 < len(blocks) and torch.is_grad_enabled()
                        and getattr(blk, "_fused_act", None) == "relu" and blk.relu_chain_ok(consumer=relu_in)
                        and blocks[i + 1].relu_chain_ok())
            if relu_in or delegate:
                out = blk(out, input_is_relu=relu_in, premasked=delegate)
            else:
                out = blk(out)
            relu_in = delegate
        return out

    def forward(self, x, cat_slot=False):
        """``cat_slot=True`` (an extension used by ``UNetSpherical.encode``): the output is written as the second half of
        a ``[B, V, 2C]`` buffer, ready for the decoder's skip concatenation (``layers.unpool_cat``)."""
        one_launch = (self.rezero and self._fused_tail is not None and _LINEAR_REZERO
                      and hasattr(self.res_connection, "forward_rezero"))
        if one_launch and self._fork is not None and self.res_connection.fusable_rezero(x):
            # Linear skip + scale + add in one launch; the block's input gradient (convolution branch + skip) is finished
            # by the fork node instead of an element-wise add
            state = None
            out = x
            if torch.is_grad_enabled() and x.requires_grad:
                out, state = self._fork(x)
            out = self._run_convs(out)
            return self.res_connection.forward_rezero(x, out, self.rezero_weight, cat_slot=cat_slot, fork_state=state)
        out = self._run_convs(x)
        if self.rezero and self._fused_tail is not None:
            if one_launch:
                return self.res_connection.forward_rezero(x, out, self.rezero_weight, cat_slot=cat_slot and _FUSED_SKIPS)
            return self._fused_tail(out, self.res_connection(x), self.rezero_weight)
        if self.rezero:
            out *= self.rezero_weight
        out += self.res_connection(x)
        return out


def _coarsen(sampling: str, kwargs: Dict, factor: int) -> Dict:
    """reference ``utils_models.py:91-102``."""
    new = dict(kwargs)
    if sampling == "equiangular":
        new["nlat"] //= factor
        new["nlon"] //= factor
    elif sampling in ("healpix", "icosahedral", "cubed"):
        new["subdivisions"] //= factor
    elif sampling == "gauss":
        new["nlat"] //= factor
    return new


def _graph_xyz(sampling: str, kwargs: Dict) -> np.ndarray:
    if sampling == "healpix":
        if not kwargs.get("nest", True):
            raise NotImplementedError("only nested HEALPix ordering is synthesised")
        return G.healpix_nested_xyz(kwargs["subdivisions"])
    if sampling == "equiangular":
        return G.equiangular_xyz(kwargs["nlat"], kwargs["nlon"])
    raise NotImplementedError(f"sampling '{sampling}' needs pygsp, which is outside the hot path")


class UNetSpherical(torch.nn.Module):
    """3-level spherical U-Net with residual blocks (reference ``my_models_graph.py:220-564``).

    ``forward(x[sample, time, node, feature]) -> [sample, time_out, node, feature_out]`` for the
    default ``dim_order``.  Extra keyword arguments not in the reference:

    * ``laplacians`` — three prepared sparse-COO Laplacians to use instead of building graphs
      (parity tests hand over the reference's own tensors);
    * ``pool_matrices`` — ``[(pool, unpool), (pool, unpool)]`` scipy matrices for the general pools;
    * ``backend`` — layer classes (see module docstring).
    """

    def __init__(self, tensor_info: Dict, sampling: str, sampling_kwargs: Dict, kernel_size_conv: int = 3,
                 conv_type: str = "graph", graph_type: str = "knn", knn: int = 20, periodic_padding: bool = True,
                 bias: bool = True, batch_norm: bool = False, batch_norm_before_activation: bool = False,
                 activation: bool = True, activation_fun: str = "relu", pool_method: str = "max",
                 kernel_size_pooling: int = 4, skip_connection: str = "stack", increment_learning: bool = False,
                 laplacians: Optional[Sequence[torch.Tensor]] = None, pool_matrices=None, backend=None):
        super().__init__()
        backend = backend or _default_backend()
        self.dim_names = tensor_info["dim_order"]["dynamic"]
        self.input_n_feature = tensor_info["input_n_feature"]
        self.output_n_feature = tensor_info["output_n_feature"]
        self.input_n_time = tensor_info["input_n_time"]
        self.output_n_time = tensor_info["output_n_time"]
        self.input_n_node = tensor_info["input_shape_info"]["dynamic"]["node"]
        self.output_n_node = tensor_info["output_shape_info"]["dynamic"]["node"]
        self.input_channels = self.input_n_feature * self.input_n_time
        self.output_channels = self.output_n_feature * self.output_n_time
        self.increment_learning = increment_learning

        sampling = sampling.lower()
        pool_method = pool_method.lower()
        conv_type = conv_type.lower()
        if conv_type not in ("graph", "image"):
            raise ValueError("{} conv_type is not supported. Choose either 'graph' or 'image'".format(conv_type))
        if conv_type == "image" and sampling != "equiangular":
            raise ValueError("conv_type='image' needs the equiangular sampling")
        if graph_type != "knn" and laplacians is None:
            raise NotImplementedError("voronoi (cotan) Laplacians need igl; pass laplacians=[...] instead")
        if skip_connection not in ("none", "stack", "sum", "avg", None):
            raise ValueError("'skip_connection' must be one of ('none', 'stack', 'sum', 'avg')")

        lonlat_ratio = None
        if sampling == "equiangular":  # reference my_models_graph.py:376-384
            lonlat_ratio = sampling_kwargs["nlon"] // sampling_kwargs["nlat"]
        cb_kwargs = dict(kernel_size=kernel_size_conv, conv_type=conv_type, bias=bias, batch_norm=batch_norm,
                         batch_norm_before_activation=batch_norm_before_activation, activation=activation,
                         activation_fun=activation_fun, periodic_padding=periodic_padding, lonlat_ratio=lonlat_ratio,
                         backend=backend)

        depth = 3  # hard-coded in the reference (my_models_graph.py:374)
        factor = int(np.sqrt(kernel_size_pooling))
        level_kwargs = [dict(sampling_kwargs, k=knn)]
        for _ in range(1, depth):
            level_kwargs.append(_coarsen(sampling, level_kwargs[-1], factor))
        if laplacians is None and conv_type == "image":
            # the image convolution has no use for a graph: identity operators keep the attribute and the buffers' shapes
            sizes = [kw["nlat"] * kw["nlon"] for kw in level_kwargs]
            laplacians = [G.scipy_to_torch_coo(__import__("scipy.sparse", fromlist=["identity"]).identity(n, format="coo")) for n in sizes]
        if laplacians is None:
            laplacians = [
                G.prepare_torch_laplacian(G.knn_laplacian(_graph_xyz(sampling, kw), knn)) for kw in level_kwargs
            ]
        if len(laplacians) != depth:
            raise ValueError(f"need {depth} laplacians")
        self.laplacians = list(laplacians)
        n_nodes = [lap.shape[0] for lap in self.laplacians]

        if pool_method in ("interp", "maxval", "maxarea", "learn"):
            if pool_matrices is None:
                if sampling != "healpix":
                    raise NotImplementedError("general pools need pool_matrices= for non-nested samplings")
                pool_matrices = [G.nested_pool_matrices(n_nodes[i], kernel_size_pooling) for i in range(depth - 1)]
            self.pool1, self.unpool1 = backend.general_pools(pool_method=pool_method, matrices=pool_matrices[0])
            self.pool2, self.unpool2 = backend.general_pools(pool_method=pool_method, matrices=pool_matrices[1])
        elif pool_method in ("max", "avg"):
            if sampling == "healpix":
                pool_cls, unpool_cls = backend.healpix_pools[pool_method]
                pkw = dict(kernel_size=kernel_size_pooling)
            elif sampling == "equiangular":  # dense 2-D index pools (torch ops)
                from .layers_equiangular import EQUIANGULAR_POOL

                pool_cls, unpool_cls = EQUIANGULAR_POOL[pool_method]
                pkw = dict(kernel_size=kernel_size_pooling, lonlat_ratio=lonlat_ratio)
            else:
                raise NotImplementedError("index pools exist for the nested HEALPix and the equiangular samplings only")
            self.pool1, self.unpool1 = pool_cls(**pkw), unpool_cls(**pkw)
            self.pool2, self.unpool2 = pool_cls(**pkw), unpool_cls(**pkw)
        else:
            raise ValueError("Not valid pooling method provided.")

        fused_skips = _FUSED_SKIPS and conv_type == "graph" and skip_connection == "stack"
        self._pool_fork = getattr(backend, "pool_fork", None) if fused_skips else None
        self._unpool_cat = getattr(backend, "unpool_cat", None) if fused_skips else None
        L0, L1, L2 = self.laplacians
        self.conv1 = ResBlock(self.input_channels, (64, 128), laplacian=L0, convblock_kwargs=cb_kwargs)
        self.conv2 = ResBlock(128, (192, 256), laplacian=L1, convblock_kwargs=cb_kwargs)
        self.conv3 = ResBlock(256, (512, 256), laplacian=L2, convblock_kwargs=cb_kwargs)
        self.uconv2 = ResBlock(512, (256, 128), laplacian=L1, convblock_kwargs=cb_kwargs)
        self.uconv1 = ResBlock(256, (128, 64), laplacian=L0, convblock_kwargs=cb_kwargs)
        self.uconv1_final = ResBlock(64, self.output_channels, laplacian=L0, convblock_kwargs=cb_kwargs)
        if self.increment_learning:
            self.res_increment = torch.nn.Parameter(torch.zeros(1), requires_grad=True)

    # reference: my_models_graph.py:492-525
    def encode(self, x):
        batch = x.shape[0]
        x_last = x[:, -1, :, -2:].unsqueeze(dim=1)
        x = x.rename(*self.dim_names).align_to("sample", "node", "time", "feature").rename(None)
        x = x.reshape(batch, self.input_n_node, self.input_channels)
        if self._pool_fork is not None:  # skip tensors land in their concatenation buffers; pool + skip share one node
            enc1 = self.conv1(x, cat_slot=True)
            pooled1, idx1, enc1 = self._pool_fork(self.pool1, enc1)
            enc2 = self.conv2(pooled1, cat_slot=True)
            pooled2, idx2, enc2 = self._pool_fork(self.pool2, enc2)
            enc3 = self.conv3(pooled2)
            return enc3, enc2, enc1, idx2, idx1, x_last
        enc1 = self.conv1(x)
        pooled1, idx1 = self.pool1(enc1)
        enc2 = self.conv2(pooled1)
        pooled2, idx2 = self.pool2(enc2)
        enc3 = self.conv3(pooled2)
        return enc3, enc2, enc1, idx2, idx1, x_last

    # reference: my_models_graph.py:528-564
    def decode(self, enc3, enc2, enc1, idx2, idx1, x_last):
        if self._unpool_cat is not None:
            x = self.uconv2(self._unpool_cat(self.unpool2, enc3, idx2, enc2))
            x = self.uconv1(self._unpool_cat(self.unpool1, x, idx1, enc1))
        else:
            x = self.unpool2(enc3, idx2)
            x = self.uconv2(torch.cat((x, enc2), dim=2))
            x = self.unpool1(x, idx1)
            x = self.uconv1(torch.cat((x, enc1), dim=2))
        x = self.uconv1_final(x)
        batch = x.shape[0]
        x = x.reshape(batch, self.output_n_node, self.output_n_time, self.output_n_feature)
        x = x.rename("sample", "node", "time", "feature").align_to(*self.dim_names).rename(None)
        if self.increment_learning:
            x *= self.res_increment
            x += x_last
        return x

    def forward(self, x):
        return self.decode(*self.encode(x))


def default_tensor_info(n_nodes: int, input_n_feature: int = 7, output_n_feature: int = 2, input_n_time: int = 3,
                        output_n_time: int = 1) -> Dict:
    """The ``tensor_info`` dict the reference's training script hands to the model
    (``my_models_graph.py:318-332``), for the 7-variable / 3-timestep configuration."""
    return {
        "dim_order": {"dynamic": ["sample", "time", "node", "feature"]},
        "input_n_feature": input_n_feature, "output_n_feature": output_n_feature,
        "input_n_time": input_n_time, "output_n_time": output_n_time,
        "input_shape_info": {"dynamic": {"node": n_nodes}},
        "output_shape_info": {"dynamic": {"node": n_nodes}},
    }


def deterministic_fill(model: torch.nn.Module, seed: int = 0, rezero: float = 1.0) -> None:
    """Deterministic, torch-version-independent parameter fill used by benchmarks and parity
    harnesses: one numpy PCG64 stream per parameter name (He-normal for ConvCheb weights,
    1/sqrt(fan_in) for the linear skips, N(0, 0.1) biases); every ``rezero_weight`` /
    ``res_increment`` is set to ``rezero`` — they initialise to 0 (``my_models_graph.py:193``),
    which would switch every convolution branch off."""
    import zlib

    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("rezero_weight") or name.endswith("res_increment"):
                p.fill_(rezero)
                continue
            rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
            if p.dim() == 3:  # ConvCheb weight [Fin, K, Fout]
                std = float(np.sqrt(2.0 / (p.shape[0] * p.shape[1])))
            elif p.dim() == 2:  # Linear skip [out, in]
                std = 1.0 / np.sqrt(p.shape[1])
            else:
                std = 0.1
            p.copy_(torch.from_numpy((rng.standard_normal(tuple(p.shape)) * std).astype(np.float32)))
