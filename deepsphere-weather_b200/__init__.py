"""deepsphere-weather_b200 — B200 (sm_100a) implementation of DeepSphere-Weather's spherical graph
convolution hot path (Chebyshev SpMM recurrence + channel mix, sparse pooling / unpooling) behind
the reference's ``modules/layers.py`` interface.  Import as ``deepsphere_weather_b200``.

Submodules: ``layers`` (drop-in modules), ``models`` (UNetSpherical and blocks), ``functional``
(autograd functions over the C-ABI), ``graphs`` (host-side operator construction), ``ddp``
(batch-sharded data parallelism), ``build`` (nvcc recipe for ``libdsw.so``).
"""
__version__ = "0.1.0"
__all__ = ["layers", "models", "functional", "graphs", "ddp", "build"]
