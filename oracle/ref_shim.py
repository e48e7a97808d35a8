"""TEST INFRASTRUCTURE ONLY -- import shim for the *real* reference.

Makes the reference's own hot-path modules importable from ``/root/reference``
in the build container, or from the byte-for-byte staged copy ``oracle/_ref`` (``oracle/stage_ref.py``,
git-ignored, travels with the gpurun snapshot) on the GPU box.  Used by ``bench.py``'s reference arm, by
``oracle/make_golden.py`` to generate the fixtures under ``tests/golden/`` and by
the CPU tests that cross-check ``oracle/restate.py`` when the reference tree is
present.  Nothing in the product package imports this file.

Shims:
  * stub ``matplotlib`` / ``matplotlib.pyplot`` (``federated_learning/utils.py:21-22``);
  * ``numpy.math = math`` (``utils_shapley.py:190`` uses ``np.math.factorial``);
  * stub ``wolframclient`` and put ``/root/reference/shapleyserver`` on ``sys.path``
    so ``compared_methods.py:3-4, 9`` imports.
"""
from __future__ import annotations

import contextlib
import io
import math
import os
import sys
import types

STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # written by oracle/stage_ref.py


def _reference_root() -> str:
    """$SVIT_REFERENCE_ROOT, else /root/reference (build container), else the staged copy oracle/_ref (GPU box)."""
    env = os.environ.get("SVIT_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/shapleyserver/fed_client_contribution"):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _reference_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "shapleyserver", "fed_client_contribution"))


_loaded = None


def load():
    """Returns a namespace with the reference's hot-path symbols."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import numpy as np

    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    plt.switch_backend = lambda *a, **k: None
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    if not hasattr(np, "math"):
        np.math = math
    for name in ("wolframclient", "wolframclient.language", "wolframclient.evaluation"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.wlexpr = lambda s: s
            m.WolframLanguageSession = m.SecuredAuthenticationKey = m.WolframCloudSession = object
            sys.modules[name] = m
    # The product repo ships its own drop-in ``shapleyserver`` package, so the
    # reference's package has to be imported under a different top-level name.
    import importlib.util

    def _load_pkg(alias: str, path: str):
        spec = importlib.util.spec_from_file_location(
            alias, os.path.join(path, "__init__.py"), submodule_search_locations=[path])
        mod = importlib.util.module_from_spec(spec) if spec and spec.loader else types.ModuleType(alias)
        mod.__path__ = [path]
        sys.modules[alias] = mod
        return mod

    base = os.path.join(REFERENCE_ROOT, "shapleyserver")
    _load_pkg("refshapleyserver", base)
    for sub in ("fed_client_contribution", "federated_learning"):
        pkg = _load_pkg(f"refshapleyserver.{sub}", os.path.join(base, sub))
        # compared_methods.py:9 imports ``fed_client_contribution`` as a top-level name
        sys.modules.setdefault(sub, pkg)

    import importlib

    with contextlib.redirect_stdout(io.StringIO()):
        # federated_learning/utils.py:26 imports its sibling by the absolute name
        # ``shapleyserver.federated_learning.networks``; serve it the reference's file.
        nets = importlib.import_module("refshapleyserver.federated_learning.networks")
        sys.modules["shapleyserver.federated_learning.networks"] = nets
        game = importlib.import_module("refshapleyserver.fed_client_contribution.game")
        ush = importlib.import_module("refshapleyserver.fed_client_contribution.utils_shapley")
        sys.modules.setdefault("fed_client_contribution.utils_shapley", ush)
        cmp_ = importlib.import_module("refshapleyserver.fed_client_contribution.compared_methods")
        futils = importlib.import_module("refshapleyserver.federated_learning.utils")
        server2 = importlib.import_module("refshapleyserver.federated_learning.server2")
        client2 = importlib.import_module("refshapleyserver.federated_learning.client2")

    ns = types.SimpleNamespace(
        Game=game.Game, utils_shapley=ush, compared_methods=cmp_, fl_utils=futils,
        ServerBase=server2.ServerBase, ClientBase=client2.ClientBase,
        evaluation=futils.evaluation,
        get_aggregated_model=futils.get_aggregated_model,
        get_difference_between_network_weights=futils.get_difference_between_network_weights,
    )
    _loaded = ns
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints per batch / per evaluation; silence it."""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def seeded_randomstate_factory(seed: int):
    """Replacement for ``np.random.RandomState`` inside ``utils_shapley`` whose
    ``RandomState(None)`` calls (``utils_shapley.py:253, 278``) are unseedable."""
    import numpy as np

    real = np.random.RandomState

    def factory(arg=None):
        return real(seed if arg is None else arg)

    return factory
