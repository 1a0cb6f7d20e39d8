"""``threatengage.rl_framework`` is where most apps under apps/threatengage_runner still import the RL pipeline from
(e.g. stage03/experiments/02/bo_exp02_vFinal_home_office_app.py:20-26); at the reference's HEAD the package lives at
``core.rl_framework``.  Every ``threatengage.rl_framework[.x.y]`` import is served the SAME module object as
``core.rl_framework[.x.y]`` (compat's modules first, then the reference's own when its src/ is on sys.path), so there is
one ``ReinforcementLearningPipeline`` class whichever path a script uses."""
import importlib
import importlib.abc
import importlib.machinery
import sys

_OLD, _NEW = __name__, "core.rl_framework"


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(_OLD + "."):
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(_NEW + spec.name[len(_OLD):])

    def exec_module(self, module):
        pass


if not any(type(f).__name__ == "_AliasFinder" and getattr(f, "_old", None) == _OLD for f in sys.meta_path):
    _f = _AliasFinder()
    _f._old = _OLD
    sys.meta_path.insert(0, _f)
