"""Import the REAL reference (read-only, /root/reference) for pinning the oracle.

Only usable in the build container: /root/reference does not exist on the GPU
box, so every caller must go through ``have_reference()`` and skip otherwise.
The reference's ``modules.py`` star-imports ``utils.py`` which pulls plotting /
download / metrics packages that are absent here; empty stand-ins are enough
because the hot path never touches them.
"""
import os
import sys
import types

REF_SRC = "/root/reference/src"


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "modules.py"))


def load_reference_modules():
    if "modules" in sys.modules and getattr(sys.modules["modules"], "__file__", "").startswith(REF_SRC):
        return sys.modules["modules"]
    for name in ("matplotlib", "matplotlib.pyplot", "wget"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")

        class Metric:  # noqa: D401 - placeholder base class
            def __init__(self, *a, **k):
                pass

        tm.Metric = Metric
        sys.modules["torchmetrics"] = tm
    sys.path.insert(0, REF_SRC)
    try:
        import modules  # noqa: E402  (the reference's src/modules.py)
    finally:
        sys.path.remove(REF_SRC)
    return modules
