"""Import shim: the package directory is `marl-delivery_b200/` (the name the project mandates), which
is not a Python identifier.  `import marl_delivery_b200` loads that directory as this module."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "marl-delivery_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
