"""The reference itself, executed: /root/reference/dist/Zlib-main.js run UNMODIFIED under oracle/minijs
(a JavaScript interpreter written for this purpose, C++, oracle/minijs/minijs.cpp).

TEST INFRASTRUCTURE ONLY: tests/ and tests/golden/make_refjs_vectors.py use it to pin the C oracle
(oracle/zts_oracle.c) and the golden vectors to the real reference. Nothing under zlib.ts_b200/ and no timed path of
bench.py imports this. /root/reference does not exist on the GPU box: `available()` is False there and the tests fall
back to the committed vectors the interpreter produced here (tests/golden/refjs_vectors.json).
"""
import os
import shutil
import subprocess
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
BUNDLE = "/root/reference/dist/Zlib-main.js"
BIN = os.path.join(_HERE, "_ref", "minijs")
SRC = os.path.join(_HERE, "minijs", "minijs.cpp")
HARNESS = os.path.join(_HERE, "minijs", "harness.js")


def build(force=False):
    """g++ oracle/minijs/minijs.cpp -> oracle/_ref/minijs (git-ignored build product)."""
    if not force and os.path.exists(BIN) and os.path.getmtime(BIN) >= os.path.getmtime(SRC):
        return BIN
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    tmp = BIN + ".tmp%d" % os.getpid()
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-trigraphs", "-o", tmp, SRC], check=True)
    os.replace(tmp, BIN)
    return BIN


def available():
    """True where the reference sources are present (this container), False on the GPU box."""
    return os.path.exists(BUNDLE) and shutil.which("g++") is not None


class RefError(Exception):
    """An exception thrown by the reference (message = the thrown Error's message or the thrown string)."""


class Batch:
    """Collects commands for one interpreter run (the bundle is parsed once per run)."""

    def __init__(self):
        self.dir = tempfile.mkdtemp(prefix="refjs_")
        self.lines = []
        self.outs = []   # per command: output path (or None)
        self.n_files = 0

    def _file(self, data=None):
        path = os.path.join(self.dir, "f%d.bin" % self.n_files)
        self.n_files += 1
        if data is not None:
            with open(path, "wb") as f:
                f.write(bytes(data))
        return path

    def add(self, op, data, *args):
        """op in rawdeflate / rawinflate / deflate / inflate / gzip / gunzip / lengths: input bytes -> output bytes."""
        src, dst = self._file(data), self._file()
        self.lines.append(" ".join([op, src, dst] + [str(a) for a in args]))
        self.outs.append(dst)
        return len(self.lines) - 1

    def add_value(self, op, data):
        """crc32 / adler32: input bytes -> number."""
        self.lines.append(" ".join([op, self._file(data)]))
        self.outs.append(None)
        return len(self.lines) - 1

    def add_zip(self, files, date_ms):
        dst = self._file()
        parts = ["zip", dst, str(int(date_ms)), str(len(files))]
        for name, data in files:
            assert " " not in name
            parts += [self._file(data), name]
        self.lines.append(" ".join(parts))
        self.outs.append(dst)
        return len(self.lines) - 1

    def add_unzip(self, archive, verify=True):
        prefix = self._file() + "_e"
        self.lines.append(" ".join(["unzip", self._file(archive), prefix, "1" if verify else "0"]))
        self.outs.append(prefix)
        return len(self.lines) - 1

    def run(self, timeout=3600):
        """-> list of (info dict, output bytes | None) or RefError instances, one per command."""
        build()
        manifest = os.path.join(self.dir, "manifest.txt")
        with open(manifest, "w") as f:
            f.write("\n".join(self.lines) + "\n")
        p = subprocess.run([BIN, BUNDLE, HARNESS, "--", manifest], capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("minijs failed: " + p.stderr[-2000:])
        lines = [l for l in p.stdout.splitlines() if l.startswith("OK ") or l.startswith("ERR ") or l in ("OK", "ERR")]
        if len(lines) != len(self.lines):
            raise RuntimeError("minijs printed %d result lines for %d commands:\n%s" % (len(lines), len(self.lines), p.stdout[-2000:]))
        results = []
        for line, out, cmd in zip(lines, self.outs, self.lines):
            if line.startswith("ERR"):
                results.append(RefError(line[4:]))
                continue
            info = {}
            for kv in line[3:].split(" "):
                if "=" in kv:
                    k, v = kv.split("=", 1)
                    info[k] = v
            data = None
            if cmd.startswith("unzip"):
                data = []
                for i in range(int(info["n"])):
                    with open(out + str(i), "rb") as f:
                        data.append(f.read())
            elif out is not None and os.path.exists(out):
                with open(out, "rb") as f:
                    data = f.read()
            results.append((info, data))
        shutil.rmtree(self.dir, ignore_errors=True)
        return results


def _one(op, data, *args):
    b = Batch()
    b.add(op, data, *args)
    r = b.run()[0]
    if isinstance(r, RefError):
        raise r
    return r


def raw_deflate(data, compression_type=2, lazy=0):
    """new RawDeflate(data, {compressionType, lazy}).compress() as the reference computes it."""
    return _one("rawdeflate", data, compression_type, lazy)[1]


def raw_inflate(stream, index=0, buffer_type=1):
    """-> (output, ip) of new RawInflate(stream, {index, bufferType}).decompress()."""
    info, out = _one("rawinflate", stream, index, buffer_type)
    return out, int(info["ip"])
