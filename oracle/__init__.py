"""Python loader of the CPU oracle (oracle/zts_oracle.c). TEST INFRASTRUCTURE ONLY:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Parity: pinned to the reference executed under oracle/minijs (tests/test_refjs.py, tests/golden/refjs_vectors.json;
see zts_oracle.h)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzts_oracle.so")

NONE, FIXED, DYNAMIC = 0, 1, 2
ERR_TEXT = {
    1: "input buffer is broken", 2: "unknown BTYPE: 3", 3: "invalid code length", 4: "output overflow",
    5: "invalid uncompressed block header: LEN", 6: "invalid compression type", 7: "out of memory",
    8: "undefined behaviour in the reference (incomplete code / distance before start)",
}

_lib = None


def build():
    """The C restatement, and -- where the reference is present (this container, not the GPU box) -- the interpreter
    that executes it (oracle/_ref/minijs)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    try:
        from . import refjs
        if refjs.available():
            refjs.build()
    except Exception:  # the interpreter only serves the CPU tests; they say so themselves if it is missing
        pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        vp, sz, u32, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int
        L.zo_crc32_update.argtypes = [vp, sz, u32]
        L.zo_crc32_update.restype = u32
        L.zo_adler32_update.argtypes = [u32, vp, sz]
        L.zo_adler32_update.restype = u32
        L.zo_lz77_encode.argtypes = [vp, sz, i32, vp, ctypes.POINTER(sz), vp, vp]
        L.zo_lz77_encode.restype = i32
        L.zo_get_lengths.argtypes = [vp, i32, i32, vp]
        L.zo_get_lengths.restype = i32
        L.zo_raw_deflate.argtypes = [vp, sz, i32, i32, vp, sz, sz, ctypes.POINTER(sz)]
        L.zo_raw_deflate.restype = i32
        L.zo_raw_deflate_dict.argtypes = [vp, sz, sz, i32, i32, vp, sz, ctypes.POINTER(sz)]
        L.zo_raw_deflate_dict.restype = i32
        L.zo_raw_deflate_bound.argtypes = [sz]
        L.zo_raw_deflate_bound.restype = sz
        L.zo_raw_inflate.argtypes = [vp, sz, sz, vp, sz, ctypes.POINTER(sz), ctypes.POINTER(sz), i32]
        L.zo_raw_inflate.restype = i32
        L.zo_diag_rpm_freq_oob.restype = ctypes.c_uint64
        L.zo_deflate_chunks_mt.argtypes = [vp, sz, sz, i32, i32, ctypes.POINTER(ctypes.c_uint64)]
        L.zo_deflate_chunks_mt.restype = i32
        L.zo_inflate_streams_mt.argtypes = [vp, vp, vp, sz, vp, sz, i32, ctypes.POINTER(ctypes.c_uint64)]
        L.zo_inflate_streams_mt.restype = i32
        L.zo_zlib_chunks_keep_mt.argtypes = [vp, sz, sz, i32, vp, sz, vp]
        L.zo_zlib_chunks_keep_mt.restype = i32
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code):
        detail = code >> 8   # the N of 'invalid code length: N'; 1 = NLEN for the uncompressed block header
        code &= 0xFF
        text = ERR_TEXT.get(code, f"oracle error {code}")
        if code == 3 and detail:
            text += ": %d" % detail
        if code == 5 and detail:
            text = "invalid uncompressed block header: NLEN"
        super().__init__(text)
        self.code = code
        self.detail = detail


def _u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def crc32(data, crc=0):
    a = _u8(data)
    return int(lib().zo_crc32_update(a.ctypes.data, a.size, crc))


def adler32(data, adler=1):
    a = _u8(data)
    return int(lib().zo_adler32_update(adler, a.ctypes.data, a.size))


def raw_deflate(data, compression_type=DYNAMIC, lazy=0, prefix=b""):
    """new RawDeflate(data, {compressionType, lazy, outputBuffer: prefix.., outputIndex}).compress()"""
    a = _u8(data)
    cap = int(lib().zo_raw_deflate_bound(a.size)) + len(prefix)
    out = np.zeros(cap, dtype=np.uint8)
    out[:len(prefix)] = np.frombuffer(bytes(prefix), dtype=np.uint8)
    n = ctypes.c_size_t(0)
    rc = lib().zo_raw_deflate(a.ctypes.data, a.size, compression_type, lazy, out.ctypes.data, cap, len(prefix),
                              ctypes.byref(n))
    if rc:
        raise OracleError(rc)
    return out[:n.value].tobytes()


def raw_deflate_dict(data, dict_len, bfinal=True, compression_type=DYNAMIC):
    """One block over data[dict_len:] with data[:dict_len] as LZ77 history (zo_raw_deflate_dict)."""
    a = _u8(data)
    cap = int(lib().zo_raw_deflate_bound(a.size))
    out = np.zeros(cap, dtype=np.uint8)
    n = ctypes.c_size_t(0)
    rc = lib().zo_raw_deflate_dict(a.ctypes.data, a.size, dict_len, 1 if bfinal else 0, compression_type,
                                   out.ctypes.data, cap, ctypes.byref(n))
    if rc:
        raise OracleError(rc)
    return out[:n.value].tobytes()


def primed_blocks(data, chunk=32768, compression_type=DYNAMIC):
    """The blocks the engine's dictionary-primed mode writes for one item, in order: one per chunk, each searching
    the 32 KiB before it as well; all but the last with BFINAL = 0 (the engine joins them with the byte-aligning
    empty stored block of SURVEY App. A.7)."""
    a = _u8(data)
    n = a.size
    n_chunks = max(1, -(-n // chunk))
    blocks = []
    for k in range(n_chunks):
        lo, hi = k * chunk, min(n, (k + 1) * chunk)
        d = min(lo, 32768)
        blocks.append(raw_deflate_dict(a[lo - d:hi], d, k + 1 == n_chunks, compression_type))
    return blocks


def raw_inflate(data, index=0, out_cap=None, mirror_readbits_quirk=False):
    """new RawInflate(data, {index}).decompress() -> (output bytes, ip)"""
    a = _u8(data)
    cap = out_cap if out_cap is not None else max(1 << 16, a.size * 64)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    n = ctypes.c_size_t(0)
    ip = ctypes.c_size_t(0)
    rc = lib().zo_raw_inflate(a.ctypes.data, a.size, index, out.ctypes.data, cap, ctypes.byref(n), ctypes.byref(ip),
                              1 if mirror_readbits_quirk else 0)
    if rc:
        raise OracleError(rc)
    return out[:n.value].tobytes(), ip.value


def lz77(data, lazy=0):
    """LZ77.encode() -> (Uint16 token array, litlen freqs[286], dist freqs[30])"""
    a = _u8(data)
    tok = np.zeros(2 * a.size + 2, dtype=np.uint16)
    fl = np.zeros(286, dtype=np.uint32)
    fd = np.zeros(30, dtype=np.uint32)
    n = ctypes.c_size_t(0)
    rc = lib().zo_lz77_encode(a.ctypes.data, a.size, lazy, tok.ctypes.data, ctypes.byref(n), fl.ctypes.data,
                              fd.ctypes.data)
    if rc:
        raise OracleError(rc)
    return tok[:n.value], fl, fd


def get_lengths(freqs, limit):
    f = np.ascontiguousarray(freqs, dtype=np.uint32)
    out = np.zeros(f.size, dtype=np.uint8)
    rc = lib().zo_get_lengths(f.ctypes.data, f.size, limit, out.ctypes.data)
    if rc:
        raise OracleError(rc)
    return out


def deflate_chunks_mt(data, chunk=65536, compression_type=DYNAMIC, threads=1):
    """RawDeflate of every chunk on `threads` host threads (baseline timing). Returns total compressed bytes."""
    a = _u8(data)
    total = ctypes.c_uint64(0)
    rc = lib().zo_deflate_chunks_mt(a.ctypes.data, a.size, chunk, compression_type, threads, ctypes.byref(total))
    if rc:
        raise OracleError(rc)
    return int(total.value)


def zlib_chunks_keep_mt(data, chunk=65536, threads=1):
    """Zlib stream (78 9C | RawDeflate(chunk) | Adler-32 BE) of every chunk, kept: returns (slots array, slot size,
    lengths). Used to make config C3's input, "zlib streams produced by the reference"."""
    a = _u8(data)
    n_chunks = (a.size + chunk - 1) // chunk
    slot = int(lib().zo_raw_deflate_bound(chunk)) + 6
    out = np.empty(n_chunks * slot, dtype=np.uint8)
    lens = np.zeros(n_chunks, dtype=np.uint64)
    rc = lib().zo_zlib_chunks_keep_mt(a.ctypes.data, a.size, chunk, threads, out.ctypes.data, slot, lens.ctypes.data)
    if rc:
        raise OracleError(rc)
    return out, slot, lens


def inflate_streams_mt(comp, offs, lens, out_stride, threads=1):
    """RawInflate of independent streams on `threads` host threads. Returns (output array, total bytes)."""
    c = _u8(comp)
    o = np.ascontiguousarray(offs, dtype=np.uint64)
    l = np.ascontiguousarray(lens, dtype=np.uint64)
    out = np.zeros(len(o) * out_stride, dtype=np.uint8)
    total = ctypes.c_uint64(0)
    rc = lib().zo_inflate_streams_mt(c.ctypes.data, o.ctypes.data, l.ctypes.data, len(o), out.ctypes.data, out_stride,
                                     threads, ctypes.byref(total))
    if rc:
        raise OracleError(rc)
    return out, int(total.value)


def smallest_block(data, dict_len=0, bfinal=True):
    """The block ZLB_MODE_SMALLEST writes for data[dict_len:]: the shortest of the reference's dynamic, fixed and
    stored constructions over the same tokens (stored only if strictly shorter than both, fixed only if strictly
    shorter than dynamic). Returns (bytes, 'dynamic' | 'fixed' | 'stored')."""
    dyn = raw_deflate_dict(data, dict_len, bfinal, DYNAMIC)
    fix = raw_deflate_dict(data, dict_len, bfinal, FIXED)
    sto = raw_deflate_dict(data, dict_len, bfinal, NONE)
    if len(_u8(data)) - dict_len > 0 and len(sto) < min(len(dyn), len(fix)):
        return sto, "stored"
    if len(fix) < len(dyn):
        return fix, "fixed"
    return dyn, "dynamic"
