"""Import shim: the product package lives in ``obj-gan_b200/`` (the directory name the
task layout mandates, which is not a valid Python identifier).  Importing ``objgan_b200``
re-points the package search path at that directory, so ``objgan_b200.ops`` is
``obj-gan_b200/ops.py`` and so on."""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_real = _os.path.join(_os.path.dirname(_here), "obj-gan_b200")
__path__.insert(0, _real)

with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
