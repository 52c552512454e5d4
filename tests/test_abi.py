"""The C-ABI library loads and exports every symbol include/nagp.h declares; the ctypes signature
table mirrors the header; without a GPU the entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nagp.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int32_t|int64_t|void|const char \*)\s*(nagp_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",")]
        nargs = 0 if args == ["void"] else len(args)
        out[m.group(2)] = (m.group(1), nargs)
    return out


@pytest.fixture(scope="module")
def lib():
    from nowcastautogp_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    decl = declared()
    assert len(decl) >= 16
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/nagp.h but not exported by libnagp.so"


def test_signature_table_matches_header(lib):
    from nowcastautogp_b200 import _lib
    decl = declared()
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, (ret, nargs) in decl.items():
        res, args = _lib.SIGNATURES[name]
        assert len(args) == nargs, name
        assert (res is None) == (ret == "void"), name


def test_version_and_no_gpu_behaviour(lib):
    import torch
    assert lib.nagp_version() >= 100
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    ctx = C.c_void_p()
    rc = lib.nagp_init(0, C.byref(ctx))
    assert rc == -2 and not ctx.value                       # NAGP_E_CUDA, no context
    assert b"CUDA" in lib.nagp_last_error(None)
    from nowcastautogp_b200.engine import Engine, NagpError
    with pytest.raises(NagpError):
        Engine(0)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from nowcastautogp_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libnagp.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nowcastautogp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("CPU oracle", "") or f == "synthetic.py" \
                    or "import oracle" not in text and "from oracle" not in text, f
                assert "from oracle" not in text and "import oracle" not in text, f
