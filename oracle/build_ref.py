"""Build recipe for the REFERENCE's own hash-grid CUDA extension (test infrastructure only).

Compiles /root/reference/hashencoder/src/{hashencoder.cu,bindings.cpp} *where they lie*
(no copy of reference sources into this repo) into oracle/_ref/_hash_encoder_ref.so for sm_100a.
The flags mirror the reference's JIT recipe (hashencoder/backend.py:12-24) plus an explicit
sm_100a gencode.  The resulting pybind module exposes the reference's three entry points
(hashencoder/src/bindings.cpp:5-9) and is used ONLY by tests / fixture generation on the GPU box to
pin oracle/hash_oracle.c and the product kernels against the reference's real kernels.

oracle/_ref/ is git-ignored (never in history) but travels to the GPU box with gpurun.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/hashencoder/src"
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "_hash_encoder_ref.so")


def build(force: bool = False) -> str | None:
    if os.path.exists(OUT_SO) and not force:
        return OUT_SO
    if not os.path.isdir(REF_SRC):
        return None  # GPU box: only the prebuilt file is used
    os.makedirs(os.path.join(OUT_DIR, "build"), exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    load(
        name="_hash_encoder_ref",
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=[
            "-O3", "-std=c++17", "-allow-unsupported-compiler",
            "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
            "-gencode=arch=compute_100a,code=sm_100a",
        ],
        sources=[os.path.join(REF_SRC, f) for f in ("hashencoder.cu", "bindings.cpp")],
        build_directory=os.path.join(OUT_DIR, "build"),
        verbose=False,
        is_python_module=False,
    )
    shutil.copyfile(os.path.join(OUT_DIR, "build", "_hash_encoder_ref.so"), OUT_SO)
    return OUT_SO


def load_ref():
    """Import the prebuilt reference extension (returns None when it was never built)."""
    if not os.path.exists(OUT_SO):
        return None
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location("_hash_encoder_ref", OUT_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
