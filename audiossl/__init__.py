"""``audiossl`` import paths for the B200 path: ``audiossl.X`` resolves to ``audiossl_b200.X`` (same module object),
so recipe code written against the reference - ``from audiossl.methods.atst.model import ATSTLightningModule``,
``from audiossl.transforms.byol_a import Mixup``, ``audiossl.methods.atstframe.embedding.load_model`` - and pickled
checkpoints that name those modules run on this framework unchanged.  Put the repo root on ``sys.path`` INSTEAD of the
reference checkout; parts of the reference outside the hot path (SURVEY.md section 2 "-" rows) do not exist here and
raise ``ModuleNotFoundError``.  Like the reference's ``__init__`` (audiossl/__init__.py:1-3) the host-side BLAS thread
pools are limited to one thread per process: the arithmetic runs on the GPU."""
import importlib
import importlib.abc
import importlib.util
import os
import sys

os.environ.setdefault("MKL_NUM_THREADS", "1")
os.environ.setdefault("OMP_NUM_THREADS", "1")

import audiossl_b200 as _impl  # noqa: E402

_PREFIX, _TARGET = __name__ + ".", _impl.__name__ + "."


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_PREFIX):
            return None
        real = _TARGET + fullname[len(_PREFIX):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except ModuleNotFoundError:
            return None
        return importlib.util.spec_from_loader(fullname, self, is_package=True)

    def create_module(self, spec):
        return importlib.import_module(_TARGET + spec.name[len(_PREFIX):])

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _AliasFinder())
__path__ = []  # submodules come from the finder above, never from this directory
__version__ = getattr(_impl, "__version__", "0")
