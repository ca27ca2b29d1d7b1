"""Build the native runtime in-tree:  python setup.py build_ext --inplace

Produces dear_pytorch_b200/_C.*.so (runtime, collectives, fused BN) and dear_pytorch_b200/_tc.*.so
(tcgen05 GEMMs with fused epilogues) -- sm_100a only; no other architecture is built.
The reference's counterpart is common/comm_core/setup.py:16-44 (NCCL+MPI); this
extension links neither.
"""
import glob
import os

from setuptools import setup
from torch.utils.cpp_extension import BuildExtension, CUDAExtension

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join("dear_pytorch_b200", "csrc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]
CXX_FLAGS = ["-O2", "-std=c++17", "-Wno-unused-function"]

ext = CUDAExtension(
    name="dear_pytorch_b200._C",
    sources=[os.path.join(CSRC, f) for f in
             ("bindings.cpp", "communicator.cpp", "symm_mem.cpp", "emu.cpp", "kernels.cu", "rs_pipe.cu", "bn_act.cu", "ln_fused.cu")],
    include_dirs=[os.path.join(ROOT, CSRC)],
    extra_compile_args={"cxx": CXX_FLAGS, "nvcc": NVCC_FLAGS},
    libraries=["rt"],
)



def cutlass_root():
    """CuTe/CUTLASS header tree vendored in the image (flashinfer ships CUTLASS 4.x); DEAR_CUTLASS_DIR overrides."""
    cand = [os.environ.get("DEAR_CUTLASS_DIR")]
    import importlib.util
    for pkg, sub in (("flashinfer", "data/cutlass"), ("tilelang", "3rdparty/cutlass")):
        spec = importlib.util.find_spec(pkg)
        if spec is not None and spec.submodule_search_locations:
            cand.append(os.path.join(list(spec.submodule_search_locations)[0], sub))
    for c in cand:
        if c and os.path.isfile(os.path.join(c, "include", "cutlass", "gemm", "collective", "builders", "sm100_umma_builder.inl")):
            return c
    raise RuntimeError("no CUTLASS header tree with sm_100 collectives found (set DEAR_CUTLASS_DIR)")


CUTLASS = cutlass_root()
tc_ext = CUDAExtension(
    name="dear_pytorch_b200._tc",
    # one translation unit per (operation, tile configuration): they compile in parallel
    sources=[os.path.join(CSRC, "tc_bindings.cpp"), os.path.join(CSRC, "tc_ffn_hw.cu")] + sorted(
        os.path.relpath(f, ROOT) for f in glob.glob(os.path.join(ROOT, CSRC, "tc_gemm_*.cu"))),
    include_dirs=[os.path.join(ROOT, CSRC), os.path.join(CUTLASS, "include"), os.path.join(CUTLASS, "tools", "util", "include")],
    extra_compile_args={"cxx": CXX_FLAGS, "nvcc": NVCC_FLAGS + ["--expt-extended-lambda", "-diag-suppress", "20012"]},
)

setup(
    name="dear_pytorch_b200",
    version="0.1.0",
    packages=["dear_pytorch_b200"],
    ext_modules=[ext, tc_ext],
    cmdclass={"build_ext": BuildExtension.with_options(use_ninja=True)},
)
