"""Import shim: the package directory is `openfoam-tpp_b200/` (hyphenated, as the repo
layout names it); Python cannot import a hyphenated name, so this module points its
package path at that directory and executes its __init__."""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_real = _os.path.join(_os.path.dirname(_here), "openfoam-tpp_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
