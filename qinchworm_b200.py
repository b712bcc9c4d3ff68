"""Import shim: exposes the package directory `qinchworm.jl_b200/` (whose name is not a valid
Python identifier) as the importable package `qinchworm_b200`.

    import qinchworm_b200 as qiw
    from qinchworm_b200.inchworm import inchworm
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "qinchworm.jl_b200")
__path__ = [_pkg_dir]  # makes this module a package whose submodules live in qinchworm.jl_b200/
with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
