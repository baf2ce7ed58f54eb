"""Import alias for code written against the reference package name: ``import photon.X`` resolves to ``photon_b200.X``
(the very same module objects — nothing is imported twice), e.g.::

    from photon.strategy.fedadam import FedAdam          # module paths of the reference that are merged here map too
    from photon.utils import get_parameters_from_state
    python -m photon.hydra_resolver run_uuid=demo        # == python -m photon_b200.hydra_resolver

Only names that exist in ``photon_b200`` resolve; see docs/MIGRATING.md for what differs.
"""
import importlib
import importlib.abc
import importlib.util
import sys

_SRC, _DST = __name__, "photon_b200"
# reference modules whose content lives in ONE module here: the five server optimizers share an implementation, the unigram
# metrics sit with the other language metrics (both are fed by the same fused-CE by-products)
_MOVED = {"strategy.fedadam": "strategy.strategies", "strategy.fedavg_eff": "strategy.strategies", "strategy.fedmom": "strategy.strategies",
          "strategy.fednestorov": "strategy.strategies", "strategy.fedyogi": "strategy.strategies",
          "strategy.strategy_with_cfg": "strategy.strategies", "metrics.unigram_normalized_metrics": "metrics.language",
          "conf.base_schema": "config.schema", "dataset.constants.dataset_constants_types": "dataset.dataset_types",
          "strategy.constants": "strategy.strategies"}


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, module):
        self._module = module

    def create_module(self, spec):
        return self._module

    def exec_module(self, module):   # already executed under its real name
        return None

    def get_code(self, fullname):    # lets ``python -m photon.<module>`` run the real module's code
        real = self._module.__spec__
        return real.loader.get_code(real.name)


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_SRC + "."):
            return None
        rel = fullname[len(_SRC) + 1:]
        real = f"{_DST}.{_MOVED.get(rel, rel)}"
        try:
            module = importlib.import_module(real)
        except ModuleNotFoundError as e:
            if e.name == real:
                return None
            raise
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(module), is_package=hasattr(module, "__path__"))
        return spec


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_impl = importlib.import_module(_DST)
__path__ = []          # submodules come from the finder above, never from this directory
__version__ = getattr(_impl, "__version__", "0")
