"""Importable alias of the package directory
`implementation-of-rank-algorithm-for-mainstream-recommender-systems_b200/` (its name is not a
Python identifier).  `import rank_b200` gives that package; submodules work too
(`from rank_b200.dcn import DCNModel`)."""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
_pkg = importlib.import_module(
    "implementation-of-rank-algorithm-for-mainstream-recommender-systems_b200")
sys.modules[__name__] = _pkg
# alias the submodules as well, so `rank_b200.dcn` IS the package's module (no double import)
_prefix = _pkg.__name__ + "."
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_prefix):
        sys.modules[__name__ + "." + _name[len(_prefix):]] = _mod
