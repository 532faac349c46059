"""Import shim: ``import deepsphere_weather_b200`` -> the package in ``deepsphere-weather_b200/``.

The package directory carries the upstream project's name (with a hyphen, which Python cannot
import directly); this shim extends ``__path__`` to that directory and re-exports its namespace.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "deepsphere-weather_b200")
__path__.insert(0, _pkg_dir)

with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
del _f
