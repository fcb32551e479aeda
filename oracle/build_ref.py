"""TEST INFRASTRUCTURE ONLY -- compiles the reference's own connected-components kernel
(/root/reference/sam2/csrc/connected_components.cu, unmodified, read where it lies) into
oracle/_ref/ so the GPU tests can pin our kernel and the C oracle against the real reference.

The source needs the torch extension headers (ATen / pybind11), so the recipe is
torch.utils.cpp_extension.load with TORCH_CUDA_ARCH_LIST=10.0a; outputs go to oracle/_ref/ only
(git-ignored, shipped to the GPU box by gpurun).  Runs only where /root/reference exists.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(os.environ.get("VLS_REFERENCE_ROOT", "/root/reference"), "sam2", "csrc", "connected_components.cu")
OUT = os.path.join(HERE, "_ref")
NAME = "ref_cc"


def so_path():
    if not os.path.isdir(OUT):
        return None
    for f in os.listdir(OUT):
        if f.startswith(NAME) and f.endswith(".so"):
            return os.path.join(OUT, f)
    return None


def build(verbose=False):
    if so_path() is not None:
        return so_path()
    if not os.path.exists(REF_SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load

    load(name=NAME, sources=[REF_SRC], build_directory=OUT, verbose=verbose,
         extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"])
    return so_path()


def load_ref():
    """Import the prebuilt module (GPU box: only the shipped .so is used, /root/reference is absent)."""
    p = so_path()
    if p is None:
        return None
    import importlib.util

    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
