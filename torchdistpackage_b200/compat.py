"""Run scripts written against the reference package without editing their imports.

    import torchdistpackage_b200.compat as compat
    compat.install_alias()                      # once, before the script's own imports

    from torchdistpackage import setup_distributed, tpc, NaiveDDP            # -> this package
    from torchdistpackage.parallel.pipeline_parallel.comm import send_forward
    from torchdistpackage.dist.launch_from_slurm import setup_distributed

``install_alias`` registers an import hook that maps ``torchdistpackage`` and every submodule
path of the reference (torchdistpackage/**) onto the *same module objects* of this package --
not copies, so singletons such as ``tpc`` stay single.  It is opt-in on purpose: nothing named
``torchdistpackage`` is installed by this package, an environment that also has the reference
installed keeps importing the reference until the alias is requested (the benchmark's reference
arm relies on that)."""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import sys

_IMPL = __name__.rsplit(".", 1)[0]              # "torchdistpackage_b200"


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, alias: str):
        self.alias = alias

    def _real(self, name: str):
        if name == self.alias or name.startswith(self.alias + "."):
            return _IMPL + name[len(self.alias):]
        return None

    def find_spec(self, name, path=None, target=None):
        real = self._real(name)
        if real is None:
            return None
        return importlib.machinery.ModuleSpec(name, self, origin=real)

    def create_module(self, spec):
        return importlib.import_module(spec.origin)      # the one and only module object

    def exec_module(self, module):                       # already executed under its real name
        pass


def install_alias(alias: str = "torchdistpackage", force: bool = False) -> bool:
    """Make ``import <alias>[.sub.module]`` resolve to this package.  Returns False (and does
    nothing) when a different package of that name is already imported, unless ``force``."""
    have = sys.modules.get(alias)
    impl = importlib.import_module(_IMPL)
    if have is not None and have is not impl:
        if not force:
            return False
        for k in [k for k in sys.modules if k == alias or k.startswith(alias + ".")]:
            del sys.modules[k]
    if not any(isinstance(f, _AliasFinder) and f.alias == alias for f in sys.meta_path):
        sys.meta_path.insert(0, _AliasFinder(alias))
    sys.modules[alias] = impl
    return True


def remove_alias(alias: str = "torchdistpackage") -> None:
    sys.meta_path[:] = [f for f in sys.meta_path
                        if not (isinstance(f, _AliasFinder) and f.alias == alias)]
    impl = sys.modules.get(_IMPL)
    for k in [k for k in sys.modules if k == alias or k.startswith(alias + ".")]:
        m = sys.modules[k]
        if m is impl or getattr(m, "__name__", "").startswith(_IMPL):
            del sys.modules[k]
