"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference from ``/root/reference``.

Used in the build container (where ``/root/reference`` exists) to (a) validate the restated
oracle in ``oracle/cheb_oracle.py`` and (b) generate the golden vectors under ``tests/golden/``
(``oracle/make_golden.py``).  The reference path does not exist on the GPU box, so nothing that
runs there may call :func:`load_reference`.

Two third-party imports of the reference are not installable here and are
stubbed *without touching the reference's files*:

* ``xsphere.remapping.compute_interpolation_weights`` (``modules/layers.py:16``; only *called*
  inside ``_build_interpolation_matrix``, ``layers.py:533``) -> a function that raises;
* ``pygsp.graphs.SphereHealpix`` / ``SphereEquiangular`` (``modules/utils_models.py:11-20``,
  ``modules/models.py:43-46``) -> small objects exposing ``.L``, ``.n_vertices``, ``.coords``,
  ``.signals`` built by ``deepsphere_weather_b200.graphs``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DSW_REFERENCE_ROOT", "/root/reference")
_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "layers.py"))


def _install_stubs():
    if "xsphere" not in sys.modules:
        xs = types.ModuleType("xsphere")
        xr = types.ModuleType("xsphere.remapping")

        def compute_interpolation_weights(*a, **k):  # pragma: no cover - never called
            raise RuntimeError("xsphere/CDO is not available; pass pooling matrices explicitly")

        xr.compute_interpolation_weights = compute_interpolation_weights
        xs.remapping = xr
        sys.modules["xsphere"] = xs
        sys.modules["xsphere.remapping"] = xr

    if "pygsp" not in sys.modules:
        if _REPO_ROOT not in sys.path:
            sys.path.insert(0, _REPO_ROOT)
        from deepsphere_weather_b200 import graphs as G
        import numpy as np

        class _Graph:
            def __init__(self, xyz, k):
                self.coords = xyz
                self.n_vertices = xyz.shape[0]
                self.L = G.knn_laplacian(xyz, k)
                lat = np.degrees(np.arcsin(np.clip(xyz[:, 2], -1, 1)))
                lon = np.degrees(np.arctan2(xyz[:, 1], xyz[:, 0])) % 360.0
                self.signals = {"lat": lat, "lon": lon}

        class SphereHealpix(_Graph):
            def __init__(self, subdivisions=2, nest=True, k=20, lap_type="normalized", **kw):
                assert nest, "only nested HEALPix ordering is synthesised"
                super().__init__(G.healpix_nested_xyz(subdivisions), k)

        class SphereEquiangular(_Graph):
            def __init__(self, nlat=16, nlon=32, k=20, lap_type="normalized", **kw):
                super().__init__(G.equiangular_xyz(nlat, nlon), k)

        pg = types.ModuleType("pygsp")
        pgg = types.ModuleType("pygsp.graphs")
        pgg.SphereHealpix = SphereHealpix
        pgg.SphereEquiangular = SphereEquiangular
        for missing in ("SphereIcosahedral", "SphereCubed", "SphereGaussLegendre"):
            setattr(pgg, missing, None)
        pg.graphs = pgg
        sys.modules["pygsp"] = pg
        sys.modules["pygsp.graphs"] = pgg


def load_reference():
    """Return ``(layers, my_models_graph)`` — the reference's own modules, unmodified."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    layers = importlib.import_module("modules.layers")
    models = importlib.import_module("modules.my_models_graph")
    return layers, models
