"""ORACLE — TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference).

Imports the UNMODIFIED reference Python modules on CPU:
  * stand-ins for the two absent third-party imports of modules/build.py:1-3 and
    common/type_utils.py:3 (`fvcore.common.registry.Registry`, `omegaconf.OmegaConf`) are put in
    sys.modules — no arithmetic lives there;
  * builtins.__POINTNET2_SETUP__ = True lets pointnet2_utils import without a built extension
    (pointnet2_utils.py:22-30) and `_ext` is then set to the CPU restatement (oracle.pointops_ref);
  * torch.Tensor.cuda is made the identity while a reference module runs (UnifiedSpatialCrossEncoderV2
    hard-codes .cuda(), unified_encoder.py:157,162).
Used by oracle/make_golden_model.py to generate tests/golden/model_*.npz and by the in-container
parity tests.  Nothing here can run on the GPU box (no /root/reference there).
"""
import builtins
import contextlib
import os
import sys
import types

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "modules"))


class _Registry:
    """Minimal stand-in for fvcore.common.registry.Registry (name -> object, decorator register)."""

    def __init__(self, name):
        self._name, self._obj_map = name, {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._obj_map[o.__name__] = o
                return o
            return deco
        self._obj_map[obj.__name__] = obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map


def install():
    if not available():
        raise RuntimeError("reference tree not present (GPU box?) — goldens are generated in the build container")
    if "fvcore.common.registry" not in sys.modules:
        fv, fvc, fvr = types.ModuleType("fvcore"), types.ModuleType("fvcore.common"), types.ModuleType("fvcore.common.registry")
        fvr.Registry = _Registry
        fv.common, fvc.registry = fvc, fvr
        sys.modules.update({"fvcore": fv, "fvcore.common": fvc, "fvcore.common.registry": fvr})
    if "omegaconf" not in sys.modules:
        om = types.ModuleType("omegaconf")

        class OmegaConf:
            @staticmethod
            def to_container(cfg, resolve=True):
                return dict(cfg) if cfg is not None else {}

            @staticmethod
            def create(d):
                return types.SimpleNamespace(**d)

        om.OmegaConf = OmegaConf
        sys.modules["omegaconf"] = om
    if "clip" not in sys.modules:   # model/objcls.py:6 imports it at module level (tokeniser only; never called here)
        try:
            import clip  # noqa: F401
        except ImportError:
            sys.modules["clip"] = types.ModuleType("clip")
    builtins.__POINTNET2_SETUP__ = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from oracle import pointops_ref
    import modules.third_party.pointnet2.pointnet2_modules  # noqa: F401  (appends its dir to sys.path)
    import pointnet2_utils  # the top-level name the SA modules use (pointnet2_modules.py:18-22)
    ext = pointops_ref.RefExt()
    pointnet2_utils._ext = ext
    import modules.third_party.pointnet2.pointnet2_utils as pu2
    pu2._ext = ext
    return ext


@contextlib.contextmanager
def cpu_cuda_identity():
    import torch
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def write_text_features(dirname, tensor):
    """The vision encoder loads {lang_path}/scannet_607_bert-base-uncased_id.pth (pcd_openvocab_encoder.py:46-47)."""
    import torch
    os.makedirs(dirname, exist_ok=True)
    torch.save(tensor, os.path.join(dirname, "scannet_607_bert-base-uncased_id.pth"))
    return dirname
