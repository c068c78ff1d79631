"""Import the unmodified pySOT reference from /root/reference in the build container.

pySOT/__init__.py:2 eagerly imports the strategy/controller/experimental_design
packages, which need POAP and pyDOE2 (neither installed, no network).  Empty
stand-in modules are enough for the surrogate and auxiliary_problems packages
to import; nothing from the stubs is ever executed on the hot path.

Only tools/ (golden generation) and opt-in tests use this; the GPU box has no
/root/reference.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("PYSOT_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "pySOT"))


def load(poap_module=None):
    """Returns the imported ``pySOT`` package.  ``poap_module``: optional real/stand-in POAP package."""
    if not available():
        raise ImportError("reference not present at %s" % REF_ROOT)
    if "pySOT" in sys.modules:
        return sys.modules["pySOT"]
    if poap_module is None and "poap" not in sys.modules:
        ps = types.ModuleType("poap.strategy")
        for name in ("BaseStrategy", "Proposal", "RetryStrategy"):
            setattr(ps, name, type(name, (), {}))
        pc = types.ModuleType("poap.controller")
        poap = types.ModuleType("poap")
        poap.strategy, poap.controller = ps, pc
        sys.modules.update({"poap": poap, "poap.strategy": ps, "poap.controller": pc})
    if "pyDOE2" not in sys.modules:
        sys.modules["pyDOE2"] = types.ModuleType("pyDOE2")
    import numpy as np

    if not hasattr(np, "int"):  # sop_strategy.py:457, utils.py:468 use the removed alias
        np.int = int
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import pySOT  # noqa: F401

    return sys.modules["pySOT"]
