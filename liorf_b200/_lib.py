"""Loader of liorf_b200/lib/libliorf_b200.so (the C-ABI library declared in include/liorf_b200.h)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path():
    return os.path.join(HERE, "lib", "libliorf_b200.so")


def build_library(verbose=False):
    """Compile the CUDA library for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(HERE, "csrc")] + ([] if verbose else ["-s"])
    subprocess.run(cmd, check=True)
    return library_path()


def load_library():
    """dlopen the CUDA library.  Fails loudly when it was not built — there is no fallback path."""
    global _LIB
    if _LIB is None:
        p = library_path()
        if not os.path.exists(p):
            raise RuntimeError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(liorf_b200 has no CPU fallback)")
        _LIB = C.CDLL(p)
        _LIB.liorf_version.restype = C.c_char_p
        _LIB.liorf_stream.restype = C.c_void_p
    return _LIB
