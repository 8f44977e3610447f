"""Import alias so that qibo's plugin loader finds this engine under the reference's name.

``qibo.set_backend(backend="qibotn", platform=..., runcard=...)`` imports the module ``qibotn``
and calls ``qibotn.MetaBackend.load(platform=..., runcard=...)``
(``/root/reference/src/qibotn/__init__.py:1-5``, ``backends/__init__.py:15``).  Putting the
repository root on ``sys.path`` (or installing it) therefore makes ``qibotn_b200`` the drop-in;
see INTEGRATION.md.

The reference's module names resolve too -- ``import qibotn.eval_qu``, ``from qibotn.circuit_convertor import
QiboCircuitToEinsum``, ``from qibotn.mps_utils import apply_gate, initial`` ... (what the reference's tests and
``backends/cutensornet.py`` import) -- each to the ONE module object of ``qibotn_b200`` that serves it, loaded on
first use, so caches and the library handle are not duplicated."""
import importlib
import importlib.abc
import importlib.util
import sys

from qibotn_b200 import MetaBackend, __version__  # noqa: F401
from qibotn_b200 import backends  # noqa: F401

_MODULES = {
    "qibotn.backends": "qibotn_b200.backends",
    "qibotn.backends.abstract": "qibotn_b200.backends.abstract",
    "qibotn.backends.cutensornet": "qibotn_b200.backends.cutensornet",
    "qibotn.backends.quimb": "qibotn_b200.backends.quimb",
    "qibotn.circuit_convertor": "qibotn_b200.network",          # QiboCircuitToEinsum
    "qibotn.circuit_to_mps": "qibotn_b200.mps",                 # QiboCircuitToMPS
    "qibotn.mps_utils": "qibotn_b200.mps",                      # initial, apply_gate, mps_site_right_swap
    "qibotn.mps_contraction_helper": "qibotn_b200.mps",         # MPSContractionHelper
    "qibotn.eval": "qibotn_b200.eval",
    "qibotn.eval_qu": "qibotn_b200.eval_qu",
    "qibotn.result": "qibotn_b200.result",
}


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname in _MODULES:
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        module = importlib.import_module(_MODULES[spec.name])
        self._spec = module.__spec__            # the import machinery overwrites __spec__ with the alias's
        return module

    def exec_module(self, module):
        module.__spec__ = self._spec


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
__path__ = []                                   # a package without files of its own: submodules come from the finder
