"""CPU: the C-ABI library builds, loads and exports every symbol include/zlibts_b200.h declares."""
import os
import re

import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    import zlibts_b200 as z
    lib = z.load_library()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "zlibts_b200.h")).read()
    declared = set(re.findall(r"\b(zlb_[a-z0-9_]+)\s*\(", header))
    declared -= {"zlb_ctx"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(z.EXPORTS) == declared
    assert lib.zlb_abi_version() == 1


def test_combine_helpers_match_zlib():
    import numpy as np
    import zlibts_b200 as z
    rng = np.random.default_rng(7)
    for _ in range(50):
        a = rng.integers(0, 256, rng.integers(0, 5000), dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, rng.integers(0, 70000), dtype=np.uint8).tobytes()
        assert z.crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
        assert z.adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)


def test_no_cpu_fallback():
    """Without a GPU the engine must refuse to construct (no silent CPU path)."""
    import pytest
    import torch
    import zlibts_b200 as z
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(z.EngineError):
        z.Engine(0)


def test_napi_addon_type_checks_against_the_napi_declarations():
    """napi/addon.cc cannot be built or run here (no Node, no node_api.h): it is at least type-checked, against a stub
    of the N-API declarations it uses (tests/napi_stub/node_api.h) and the real C-ABI header."""
    import shutil
    import subprocess
    import pytest
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "napi_stub"),
                        os.path.join(ROOT, "napi", "addon.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
