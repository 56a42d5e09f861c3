"""groth-sahai-rs_b200 -- B200-native (sm_100a) Groth-Sahai commit / prove / verify engine.

The package is a thin host layer over ``libgs_b200.so`` (C ABI in ``include/gs_b200.h``):

* ``ffi``   -- ctypes binding of every exported symbol (``Engine``), bytes in / bytes out
* ``api``   -- mirror of the reference's Rust API for the hot path (``CRS``, ``batch_commit_G1``,
               ``PPE(...).commit_and_prove`` / ``.verify`` ...), same names and argument meaning
* ``shard`` -- multi-GPU partition of the batch workloads (one process per GPU, verdict all-gather)

There is NO CPU fallback: importing works anywhere (so the symbol table can be checked),
but creating an ``Engine`` without the built library or without a CUDA device raises.
"""
from .ffi import Engine, GsError, lib_path, load_library, EXPORTED_SYMBOLS  # noqa: F401
from . import api  # noqa: F401
from . import shard  # noqa: F401

__all__ = ["Engine", "GsError", "lib_path", "load_library", "EXPORTED_SYMBOLS", "api", "shard"]
