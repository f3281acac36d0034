"""Compile the reference's OWN mesh-extraction natives into oracle/_ref/ (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The two Cython/C++ modules the SDF-decoder consumer uses --
``utils/libmise/mise.pyx`` (MISE octree) and ``utils/libmcubes`` (mcubes.pyx + pywrapper.cpp + marchingcubes.cpp) --
build from their own few source files: ``cython`` translates the .pyx where it lies, ``g++`` compiles the result
together with the reference's .cpp files (read in place from /root/reference, never copied into the repo).  Outputs go
to oracle/_ref/ only (git-ignored, NOT gpurun-ignored: the .so files travel to the GPU box, which has no
/root/reference).  Used by the tests as the real-reference checker for ls_mise_* / ls_mcubes_*.

    python -m oracle.build_ref
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_UTILS = os.path.join(os.environ.get("LS_REFERENCE_ROOT", "/root/reference"),
                         "lib_shape_prior/core/models/utils/occnet_utils/utils")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_UTILS, "libmise", "mise.pyx"))


def _ext(name: str) -> str:
    return os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))


def built() -> bool:
    return os.path.exists(_ext("mise")) and os.path.exists(_ext("mcubes"))


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds both modules (built now or earlier)."""
    if built() and not force:
        return True
    if not available():
        return False
    import numpy

    os.makedirs(OUT, exist_ok=True)
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include()]
    cxx = ["g++", "-O2", "-shared", "-fPIC", "-std=c++14", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]
    jobs = [
        ("mise", os.path.join(REF_UTILS, "libmise", "mise.pyx"), [], []),
        ("mcubes", os.path.join(REF_UTILS, "libmcubes", "mcubes.pyx"),
         [os.path.join(REF_UTILS, "libmcubes", f) for f in ("pywrapper.cpp", "marchingcubes.cpp")],
         # pywrapper.cpp uses numpy-1.x type macros that numpy 2 dropped: supplied on the command line, sources untouched
         ["-I" + os.path.join(REF_UTILS, "libmcubes"), "-DPyArray_DOUBLE=NPY_DOUBLE", "-DPyArray_ULONG=NPY_ULONG"]),
    ]
    for name, pyx, extra_src, extra_inc in jobs:
        gen = os.path.join(OUT, name + ".cpp")
        subprocess.run([sys.executable, "-m", "cython", "--cplus", "-3", pyx, "-o", gen], check=True)
        flags = [f for f in cxx if not (name == "mcubes" and f.startswith("-DNPY_NO_DEPRECATED"))]
        subprocess.run(flags + inc + extra_inc + [gen] + extra_src + ["-o", _ext(name)], check=True)
    return built()


def load():
    """(MISE class, marching_cubes function) of the reference, or None when oracle/_ref is not built."""
    if not built():
        return None
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import mcubes  # noqa: E402
    import mise    # noqa: E402

    return mise.MISE, mcubes.marching_cubes


if __name__ == "__main__":
    print("oracle/_ref built:", build(force="--force" in sys.argv))
