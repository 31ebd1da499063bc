"""TEST INFRASTRUCTURE -- CPU restatement of the reference's container framing, byte for byte, over the C oracle's
RawDeflate / CRC32 / Adler32 (oracle/zts_oracle.c). Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline may import this; the product path never does.

Parity status: pinned -- tests/test_refjs.py compares these functions with the bytes the reference's own Deflate /
GZip / Zip classes produce when dist/Zlib-main.js is executed under oracle/minijs (tests/golden/refjs_vectors.json);
also cross-checked in tests/ against CPython's zlib / gzip / zipfile readers.

Where the reference's own behaviour is a bug the intended bytes are restated instead, as listed in SURVEY.md
Appendix B (B-1: Deflate.compress throws a RangeError for outputs > 32 KiB; restated as header + body + Adler-32).
"""
import struct

from . import DYNAMIC, adler32, crc32, raw_deflate


def zlib_stream(data, compression_type=DYNAMIC):
    """Deflate.compress, src/Deflate.ts:60-99."""
    cmf = 120                                           # :66
    flg = (compression_type << 6) | (0 << 5)            # :74
    flg |= 31 - ((cmf << 8) + flg) % 31                 # :75-76
    body = raw_deflate(data, compression_type)          # :84-85
    return bytes([cmf, flg & 0xFF]) + body + struct.pack(">I", adler32(data))  # :81,95 (writeUintBE)


def gzip_string(s):
    """fname / fcomment, src/GZip.ts:133-150: chars > 0xFF as two bytes, NUL terminated."""
    out = bytearray()
    for ch in s:
        c = ord(ch)
        out += struct.pack("<H", c & 0xFFFF) if c > 0xFF else bytes([c])
    return bytes(out) + b"\0"


def gzip_member(data, filename="", comment="", hcrc=False, mtime=0, compression_type=DYNAMIC):
    """GZip.compress, src/GZip.ts:96-194 (MTIME is Date.now() there, :121; fixed here)."""
    flg = (0x08 if filename else 0) | (0x10 if comment else 0) | (0x02 if hcrc else 0)   # :111-117
    h = b"\x1f\x8b" + bytes([8, flg]) + struct.pack("<I", mtime & 0xFFFFFFFF) + bytes([0, 3])  # :106-128
    if filename:
        h += gzip_string(filename)
    if comment:
        h += gzip_string(comment)
    if hcrc:
        h += struct.pack("<H", crc32(h) & 0xFFFF)       # :153-156
    body = raw_deflate(data, compression_type)          # :159-166
    return h + body + struct.pack("<II", crc32(data), len(data) & 0xFFFFFFFF)  # :180-185


def dos_time(date):
    """src/Zip.ts:129-139."""
    return bytes([((date.minute & 0x7) << 5) | (date.second >> 1), (date.hour << 3) | (date.minute >> 3),
                  ((date.month & 0x7) << 5) | date.day, (((date.year - 1980) & 0x7F) << 1) | (date.month >> 3)])


def zip_archive(files, comment=b"", compression_type=DYNAMIC):
    """Zip.addFile + Zip.compress, src/Zip.ts:80-372. files: list of dicts {name, data, date, method (0 | 8),
    comment (str), os}. No extra fields (the reference writes extraFieldLength 0, :215), no encryption."""
    local, central = [], []
    offset = 0
    for f in files:
        data = bytes(f["data"])
        method = f.get("method", 8)
        body = raw_deflate(data, compression_type) if method == 8 else data   # :92-96,146-149
        name = bytes(ord(c) & 0xFF for c in f["name"])                        # stringToByteArray, :309
        fcomment = bytes(ord(c) & 0xFF for c in f.get("comment", ""))
        common = struct.pack("<HHH", 20, 0, method) + dos_time(f["date"]) + struct.pack(
            "<IIIHH", crc32(data), len(body), len(data), len(name), 0)        # :242-284
        local.append(b"PK\x03\x04" + common + name + body)                    # :228,310,330
        central.append(b"PK\x01\x02" + bytes([20, f.get("os", 0)]) + common +
                       struct.pack("<HHHII", len(fcomment), 0, 0, 0, offset) + name + fcomment)  # :234-326
        offset += 30 + len(name) + len(body)
    cd = b"".join(central)
    eocd = b"PK\x05\x06" + struct.pack("<HHHHIIH", 0, 0, len(files), len(files), len(cd), offset, len(comment))  # :340-364
    return b"".join(local) + cd + eocd + bytes(comment)
