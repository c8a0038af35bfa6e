"""Compile the reference's CPU ops into oracle/_ref/mega_ref_C*.so (see ref_wrapper.cpp).

Only possible where /root/reference exists (this container); the GPU box uses the prebuilt
file that travels with the snapshot. Outputs stay in oracle/_ref/ (git-ignored).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = os.environ.get("MEGA_REFERENCE", "/root/reference") + "/mega_core/csrc"
OUT = os.path.join(HERE, "_ref")
NAME = "mega_ref_C"


def built_path():
    for f in sorted(os.listdir(OUT)) if os.path.isdir(OUT) else []:
        if f.startswith(NAME) and f.endswith(".so"):
            return os.path.join(OUT, f)
    return None


def build(verbose=False):
    if built_path():
        return built_path()
    if not os.path.isdir(REF_CSRC):
        return None
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    load(name=NAME, sources=[os.path.join(HERE, "ref_wrapper.cpp")], extra_include_paths=[REF_CSRC],
         extra_cflags=["-O2", "-w"], build_directory=OUT, verbose=verbose, is_python_module=True)
    return built_path()


def load_module():
    """import the prebuilt extension (exports nms, roi_align_forward, ... like mega_core._C)"""
    path = built_path()
    if path is None:
        raise ImportError("oracle/_ref is not built; run oracle/build_ref.py where /root/reference exists")
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv))
