"""CPU: the C-ABI library loads, exports every function include/dv3_b200.h declares, and the
ctypes mirrors in _lib.py match the header's structs field for field (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dv3_b200.h")


def _strip_comments(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def header_functions():
    src = _strip_comments(open(HEADER).read())
    src = re.sub(r"typedef\s+(struct|enum)\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dv3_\w+)\s*\(", src)))


def header_structs():
    src = _strip_comments(open(HEADER).read())
    out = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"\[\d+\]", "", decl)
            names = re.findall(r"(\w+)\s*(?:,|$)", decl)
            fields += names
        out[name] = fields
    return out


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.lib()
    declared = header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in dv3_b200.h but not exported"
        assert name in pkg._lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(pkg._lib.SIGNATURES) == declared
    assert lib.dv3_version() == 2


def test_struct_mirrors_match_header(pkg):
    structs = header_structs()
    for cname, cls in pkg._lib.STRUCTS.items():
        assert cname in structs, cname
        assert [f for f, _ in cls._fields_] == structs[cname], cname


def test_no_torch_types_in_abi():
    src = open(HEADER).read()
    assert "torch" not in _strip_comments(src) and "at::" not in src
    assert 'extern "C"' in src


def test_missing_library_fails_loudly(pkg, monkeypatch):
    import pytest
    monkeypatch.setattr(pkg._lib, "_lib", None)
    monkeypatch.setattr(pkg._lib, "LIB_PATH", "/nonexistent/libdv3_b200.so")
    with pytest.raises(pkg._lib.Dv3Error, match="no .*fallback"):
        pkg._lib.lib()


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "dreamerv3-torch_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), fn
