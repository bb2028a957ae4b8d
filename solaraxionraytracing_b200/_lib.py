"""Loads libsart.so (the hand-written sm_100a CUDA library behind include/sart.h). There is no fallback: if the
library is missing or its ABI does not match this package, importing raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import abi

import os

# SART_LIB selects an experimental build variant of the same library (development only).
LIB_PATH = Path(os.environ.get("SART_LIB") or Path(__file__).resolve().parent / "libsart.so")


class SartError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsart error {code}: {msg}")
        self.code = code


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m solaraxionraytracing_b200.build` "
                          "(there is no CPU or PyTorch fallback for the ray-tracing path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in abi.SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.sart_abi_version() != abi.ABI_VERSION:
        raise ImportError("libsart.so ABI version mismatch")
    for what, struct in (("setup", abi.Setup), ("tables", abi.Tables), ("counters", abi.Counters)):
        n = getattr(lib, f"sart_sizeof_{what}")()
        if n != C.sizeof(struct):
            raise ImportError(f"sart_{what}_t layout drift: library {n} bytes, abi.py {C.sizeof(struct)} bytes")
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise SartError(rc, lib.sart_last_error().decode("utf-8", "replace"))
