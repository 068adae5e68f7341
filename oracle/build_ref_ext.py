"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by sceneverse_b200/).

Compiles the reference's own PointNet++ CUDA extension, from the sources where they lie under
/root/reference (never copied), for sm_100a into oracle/_ref/_ext_ref.so.

The reference's setup.py cannot be used (hard-coded TORCH_CUDA_ARCH_LIST with sm_37,
modules/third_party/pointnet2/setup.py:17), so this is a direct torch.utils.cpp_extension build of
_ext_src/src/*.{cpp,cu} with _ext_src/include on the include path.  oracle/_ref/ is git-ignored
but travels to the GPU box with gpurun, where tests/ use it (when present) as the ground truth for
the index ops and oracle/make_golden_gpu.py uses it to generate tests/golden/*.npz.
Run in the build container only (needs /root/reference).
"""
import glob
import os
import sys

REF_SRC = "/root/reference/modules/third_party/pointnet2/_ext_src"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def build(verbose=False):
    if not os.path.isdir(REF_SRC):
        return None  # GPU box: use the prebuilt file
    os.makedirs(OUT_DIR, exist_ok=True)
    so = os.path.join(OUT_DIR, "_ext_ref.so")
    srcs = sorted(glob.glob(os.path.join(REF_SRC, "src", "*.cpp")) + glob.glob(os.path.join(REF_SRC, "src", "*.cu")))
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load

    load(
        name="_ext_ref",
        sources=srcs,
        extra_include_paths=[os.path.join(REF_SRC, "include")],
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100a,code=sm_100a"],
        build_directory=OUT_DIR,
        is_python_module=False,  # build only; there is no GPU here to import against
        verbose=verbose,
    )
    return so


def load_prebuilt():
    """Import the prebuilt extension (GPU box or here). Returns the module or None."""
    so = os.path.join(OUT_DIR, "_ext_ref.so")
    if not os.path.exists(so):
        return None
    import importlib.util

    import torch  # noqa: F401  (libtorch symbols)

    spec = importlib.util.spec_from_file_location("_ext_ref", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
