"""The reference's equiangular *image* path: ``Conv2dEquiangular`` and the equiangular index pools
(``modules/layers.py:383-524, 601-781``), kept API-complete so that ``conv_type="image"`` / equiangular ``max`` / ``avg``
pooling models build against this package too.

This is deliberately NOT hand-written CUDA: it is a dense 2-D convolution on a regular lat x lon image, which cuDNN already
serves at its roofline — the graph path is the hot path this package exists for.  The modules below are thin compositions of
PyTorch ops with the reference's constructor signatures, attribute names and state-dict keys (``conv.weight``, ``conv.bias``);
they run wherever PyTorch runs.  Node order is the reference's: row-major (lat-major), ``[sample, node, feature]``.
"""
from __future__ import annotations

import torch
from torch.nn import functional as F


def get_nlat_nlon(n_nodes: int, lonlat_ratio):
    """Height (latitudes) and width (longitudes) of the image behind ``n_nodes`` row-major nodes
    (``layers.py:383-404``: ``lonlat_ratio = n_lon / n_lat``, with the same correction when the ratio is inexact)."""
    n_lat = int((n_nodes / lonlat_ratio) ** 0.5)
    n_lon = int((n_nodes * lonlat_ratio) ** 0.5)
    if n_lat * n_lon != n_nodes:
        if n_lat and n_nodes % n_lat == 0:
            n_lon = n_nodes // n_lat
        if n_lon and n_nodes % n_lon == 0:
            n_lat = n_nodes // n_lon
    if n_lat * n_lon != n_nodes:
        raise AssertionError(f"Unable to unpack nodes: {n_nodes}, lonlat_ratio: {lonlat_ratio}")
    return n_lat, n_lon


def _to_image(x: torch.Tensor, lonlat_ratio) -> torch.Tensor:
    """``[B, V, F] -> [B, F, lat, lon]``."""
    b, v, f = x.shape
    n_lat, n_lon = get_nlat_nlon(v, lonlat_ratio)
    return x.reshape(b, n_lat, n_lon, f).permute(0, 3, 1, 2)


def _to_nodes(img: torch.Tensor) -> torch.Tensor:
    """``[B, F, lat, lon] -> [B, V, F]``."""
    b, f, h, w = img.shape
    return img.permute(0, 2, 3, 1).reshape(b, h * w, f)


class Conv2dEquiangular(torch.nn.Module):
    """Square 2-D convolution on the lat x lon image; with ``periodic_padding`` the longitude axis wraps around (a
    cylinder), the latitude axis is zero-padded (``layers.py:429-524``)."""

    def __init__(self, in_channels, out_channels, kernel_size, lonlat_ratio, periodic_padding, bias, **kwargs):
        super().__init__()
        self.lonlat_ratio = lonlat_ratio
        self.periodic_padding = periodic_padding
        self.kernel_size = kernel_size
        self.pad_width = int((kernel_size - 1) / 2)
        self.conv = torch.nn.Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, bias=bias, **kwargs)
        torch.nn.init.xavier_uniform_(self.conv.weight)
        if bias:
            torch.nn.init.zeros_(self.conv.bias)

    def periodic_pad(self, x, width, periodic_padding=True):
        """``x`` is ``N x C x H x W``; wrap ``width`` columns around in longitude (or zero-pad them), zero-pad latitude."""
        if width == 0:
            return x
        if periodic_padding:
            x = F.pad(x, (width, width, 0, 0), mode="circular")
            return F.pad(x, (0, 0, width, width))
        return F.pad(x, (width, width, width, width))

    def forward(self, x):
        img = self.periodic_pad(_to_image(x, self.lonlat_ratio), self.pad_width, self.periodic_padding)
        return _to_nodes(self.conv(img))


class EquiangularMaxPool(torch.nn.Module):
    """``kernel_size`` is the number of pooled pixels (4 -> 2 x 2 windows), as in the reference (``layers.py:601-654``)."""

    def __init__(self, lonlat_ratio, kernel_size, return_indices=True, *args, **kwargs):
        super().__init__()
        self.lonlat_ratio = lonlat_ratio
        self.kernel_size = int(kernel_size**0.5)
        self.return_indices = return_indices

    def forward(self, x):
        img = _to_image(x, self.lonlat_ratio)
        if self.return_indices:
            y, idx = F.max_pool2d(img, self.kernel_size, return_indices=True)
            return _to_nodes(y), idx
        return _to_nodes(F.max_pool2d(img, self.kernel_size))


class EquiangularMaxUnpool(torch.nn.Module):
    def __init__(self, lonlat_ratio, kernel_size, *args, **kwargs):
        super().__init__()
        self.lonlat_ratio = lonlat_ratio
        k = int(kernel_size**0.5)
        self.kernel_size = (k, k)

    def forward(self, x, indices):
        return _to_nodes(F.max_unpool2d(_to_image(x, self.lonlat_ratio), indices, self.kernel_size))


class EquiangularAvgPool(torch.nn.Module):
    def __init__(self, lonlat_ratio, kernel_size, *args, **kwargs):
        super().__init__()
        self.lonlat_ratio = lonlat_ratio
        k = int(kernel_size**0.5)
        self.kernel_size = (k, k)

    def forward(self, x):
        return _to_nodes(F.avg_pool2d(_to_image(x, self.lonlat_ratio), self.kernel_size)), None


class EquiangularAvgUnpool(torch.nn.Module):
    def __init__(self, lonlat_ratio, kernel_size, *args, **kwargs):
        super().__init__()
        self.lonlat_ratio = lonlat_ratio
        self.kernel_size = int(kernel_size**0.5)

    def forward(self, x, *args):
        img = F.interpolate(_to_image(x, self.lonlat_ratio), scale_factor=(self.kernel_size, self.kernel_size), mode="nearest")
        return _to_nodes(img)


EQUIANGULAR_POOL = {"max": (EquiangularMaxPool, EquiangularMaxUnpool), "avg": (EquiangularAvgPool, EquiangularAvgUnpool)}
