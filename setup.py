"""Packaging for torchdistpackage_b200.

    pip install --no-build-isolation -e .        # develop in place (the usual way: the native
                                                 # extension is built in-tree next to the package)
    pip install --no-build-isolation .           # regular install

The native extension (hand-written sm_100a CUDA, ``torchdistpackage_b200/_C.so``) is compiled by
``torchdistpackage_b200/ops/_build.py`` -- nvcc for the ``.cu`` units, g++ for the one binding
unit that sees torch -- and shipped as package data together with its sources, so an installed
copy can rebuild itself.  Building needs nvcc (CUDA >= 12.8) and the torch the package will run
with; no GPU is needed to build (nvcc cross-compiles).  ``TDP_SKIP_NATIVE_BUILD=1`` packages the
Python side only (CPU / gloo use: every op has a torch fallback).
"""
import os
import sys

from setuptools import find_packages, setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))


class BuildWithNative(build_py):
    def run(self):
        if os.environ.get("TDP_SKIP_NATIVE_BUILD", "0") != "1":
            sys.path.insert(0, ROOT)
            try:
                from torchdistpackage_b200.ops._build import build as build_native
                so = build_native(verbose=True)
                print(f"[setup] native extension: {so}")
            finally:
                sys.path.pop(0)
        super().run()


def _version() -> str:
    for line in open(os.path.join(ROOT, "torchdistpackage_b200", "__init__.py")):
        if line.startswith("__version__"):
            return line.split("=")[1].strip().strip("\"'")
    return "0"


setup(
    name="torchdistpackage_b200",
    version=_version(),
    description="Blackwell (B200, sm_100a) native mixed-parallel training toolkit with the API of "
                "TorchDistPackage: DDP / ZeRO / TP+SP / 1F1B PP / MoE EP on NVSwitch kernels",
    packages=find_packages(include=["torchdistpackage_b200", "torchdistpackage_b200.*"]),
    package_data={"torchdistpackage_b200": ["_C.so", "csrc/*", "csrc/*/*", "tools/sbatch.sh"]},
    python_requires=">=3.10",
    install_requires=["torch>=2.6"],
    cmdclass={"build_py": BuildWithNative},
    zip_safe=False,
)
