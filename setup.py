"""Build the native runtime in-tree:  python setup.py build_ext --inplace

Produces dear_pytorch_b200/_C.*.so (sm_100a only; no other architecture is built).
The reference's counterpart is common/comm_core/setup.py:16-44 (NCCL+MPI); this
extension links neither.
"""
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
             ("bindings.cpp", "communicator.cpp", "symm_mem.cpp", "emu.cpp", "kernels.cu", "bn_act.cu")],
    include_dirs=[os.path.join(ROOT, CSRC)],
    extra_compile_args={"cxx": CXX_FLAGS, "nvcc": NVCC_FLAGS},
    libraries=["rt"],
)

setup(
    name="dear_pytorch_b200",
    version="0.1.0",
    packages=["dear_pytorch_b200"],
    ext_modules=[ext],
    cmdclass={"build_ext": BuildExtension.with_options(use_ninja=True)},
)
