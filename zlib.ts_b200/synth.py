"""Synthetic inputs of BASELINE.json's configs (SURVEY.md Appendix D generators, C implementation)."""
import ctypes
import os

import numpy as np

_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libzts_synth.so")
        if not os.path.exists(path):
            raise RuntimeError("libzts_synth.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = ctypes.CDLL(path)
        _lib.zts_gen_text.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32]
        _lib.zts_gen_mixed.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint32]
    return _lib


def text(n, seed, out=None):
    """text(n, seed): Zipf-word text."""
    buf = np.empty(n, dtype=np.uint8) if out is None else out
    _load().zts_gen_text(buf.ctypes.data, n, seed & 0xFFFFFFFF)
    return buf


def mixed(n, seed, seg=4096, out=None):
    """mixed(n, seed, seg): random / text / byte-run / 8-byte-record segments."""
    buf = np.empty(n, dtype=np.uint8) if out is None else out
    _load().zts_gen_mixed(buf.ctypes.data, n, seed & 0xFFFFFFFF, seg)
    return buf
