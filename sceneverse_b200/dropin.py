"""Registers the replacement under the import names the reference uses.

`install()` makes `import pointnet2._ext as _ext` (reference pointnet2_utils.py:22-23) resolve to
sceneverse_b200.pointnet2._ext without touching the reference tree.  See INTEGRATION.md.
"""
import sys


def install():
    from . import pointnet2 as pkg
    from .pointnet2 import _ext

    sys.modules.setdefault("pointnet2", pkg)
    sys.modules.setdefault("pointnet2._ext", _ext)
    return _ext
