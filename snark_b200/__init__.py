"""snark_b200 -- B200 (sm_100a) backend for the Groth16 prover hot path of arkworks-rs/snark.

The product is `libb200snark.so` (C ABI in include/b200snark.h, CUDA sources in snark_b200/csrc).
This package is the thin Python binding used by the tests and by bench.py; it holds no arithmetic
and has NO CPU fallback: importing works anywhere, but every call raises when the shared library
or an sm_100 GPU is missing.
"""
from .lib import B2SError, Backend, lib_path, load_library  # noqa: F401

__all__ = ["Backend", "B2SError", "lib_path", "load_library"]
