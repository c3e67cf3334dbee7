"""Import shims that let the reference's `networks/` package import on a modern stack.

The reference's networks import names that no longer exist (probed in SURVEY.md section 7):
`scipy.misc.imread/imsave/imresize/imshow` (networks/SPyNet/Network.py) and expect the
top-level modules `Stack` and `networks` on sys.path.  `install(reference_root)` injects
harmless stand-ins and the paths; the ops themselves come from this repo's `my_package`,
which must be importable first (it shadows the reference's own `my_package/`).
"""
import os
import sys
import types


def install(reference_root=None):
    import scipy.misc as misc  # noqa: F401  (exists as an empty-ish module on modern SciPy)

    def _gone(*_a, **_k):
        raise RuntimeError("scipy.misc image IO was removed from SciPy; not available in this shim")

    for name in ("imread", "imsave", "imresize", "imshow"):
        if not hasattr(misc, name):
            setattr(misc, name, _gone)
    pkg = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if pkg not in sys.path:
        sys.path.insert(0, pkg)  # our my_package wins over the reference's
    if reference_root and os.path.isdir(reference_root) and reference_root not in sys.path:
        sys.path.append(reference_root)  # networks/, Stack.py
    return types.SimpleNamespace(package_root=pkg, reference_root=reference_root)
