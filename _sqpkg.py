"""Import helper: the package directory is named `sketch-for-rna-seq_b200` (not a valid Python identifier), so it
is loaded under the alias `sketch_for_rna_seq_b200`.

    from _sqpkg import sqb
"""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "sketch-for-rna-seq_b200")
_ALIAS = "sketch_for_rna_seq_b200"


def load():
    if _ALIAS in sys.modules:
        return sys.modules[_ALIAS]
    spec = importlib.util.spec_from_file_location(
        _ALIAS, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod


sqb = load()
