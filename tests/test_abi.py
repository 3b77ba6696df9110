"""CPU tests of the drop-in boundary: the C-ABI library loads here (no GPU needed) and exports every
symbol include/neko_top_b200.h declares; argument errors come back as status codes with a message;
with no CUDA device the library refuses to create a handle (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import neko_top_b200  # noqa: F401
from neko_top_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "neko_top_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def L():
    _lib.build()
    return _lib.lib()


def test_exports_match_header(L):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS) == names, "loader table and header disagree"


def test_version_and_error_reporting(L):
    assert L.b200_version() >= 100
    _lib.set_abort_on_error(0)
    try:
        h = C.c_void_p()
        rc = L.b200_adjrhs_create(C.byref(h), C.byref(C.c_int(3)), C.byref(C.c_int(8)), C.byref(C.c_int(0)))
        assert rc == 1 and b"lx=3" in L.b200_last_error()
        rc = L.b200_adjrhs_compute(None, *([None] * 16))
        assert rc == 1
    finally:
        _lib.set_abort_on_error(1)


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _lib.set_abort_on_error(0)
    try:
        h = C.c_void_p()
        rc = L.b200_adjrhs_create(C.byref(h), C.byref(C.c_int(8)), C.byref(C.c_int(8)), C.byref(C.c_int(0)))
        assert rc == 2 and not h.value
        assert b"no CPU fallback" in L.b200_last_error()
    finally:
        _lib.set_abort_on_error(1)


def test_product_never_imports_oracle():
    """The package and bench's product arm must not reach into oracle/."""
    pkg = os.path.join(ROOT, "neko-top_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".f90")):
                src = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in src and "np_oracle" not in src and "liboracle" not in src, f


def test_fortran_shim_binds_every_symbol():
    """fortran/neko_top_b200.f90 declares a bind(c) interface for each exported entry point."""
    path = os.path.join(ROOT, "neko-top_b200", "fortran", "neko_top_b200.f90")
    src = open(path).read().lower()
    for n in _declared():
        assert f"name='{n}'" in src or f'name="{n}"' in src, f"{n} missing from the Fortran interface block"


def test_c99_consumer_links_and_runs(L):
    """csrc/host/abi_check.c (C99, -pedantic -Werror) includes the header, takes the address of every declared
    entry point, links against the shared library and exercises the no-abort error path.  No GPU needed."""
    import subprocess
    host = os.path.join(ROOT, "neko-top_b200", "csrc", "host")
    subprocess.check_call(["make", "-B", "-C", host], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(host, "abi_check")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    from neko_top_b200 import _lib
    assert f"{len(_lib.SYMBOLS)} entry points" in out.stdout


def test_fortran_shim_is_well_formed():
    """No Fortran compiler exists in the image, so the shim gets a structural lint instead: every
    subroutine / function / interface / module / type block is closed by a matching `end`, every bind(c) name
    is an entry point of the header, and continuation lines are consistent."""
    fdir = os.path.join(ROOT, "neko-top_b200", "fortran")
    for fn in sorted(os.listdir(fdir)):
        txt = open(os.path.join(fdir, fn)).read()
        code = [re.sub(r"!.*$", "", ln).strip().lower() for ln in txt.splitlines()]
        stack = []
        joined, cur = [], ""
        for ln in code:                                   # join continuation lines
            if not ln:
                continue
            cur += " " + ln.rstrip("&").lstrip("&")
            if not ln.endswith("&"):
                joined.append(cur.strip())
                cur = ""
        assert cur == "", f"{fn}: dangling continuation"
        for ln in joined:
            m = re.match(r"end\s*(subroutine|function|interface|module|type|select|if|do|associate)\b", ln)
            if m:
                kind = m.group(1)
                if kind in ("subroutine", "function", "interface", "module", "type"):
                    assert stack and stack[-1] == kind, f"{fn}: unexpected 'end {kind}' (open: {stack[-3:]})"
                    stack.pop()
                continue
            if re.match(r"(abstract\s+)?interface\b", ln):
                stack.append("interface")
            elif re.match(r"module\s+(?!procedure)\w+$", ln):
                stack.append("module")
            elif re.match(r"type\s*(,[^:]*)?::\s*\w+$", ln) or re.match(r"type\s+\w+$", ln):
                stack.append("type")
            elif re.search(r"\bsubroutine\s+\w+\s*\(", ln) and not ln.startswith("call"):
                stack.append("subroutine")
            elif re.search(r"\bfunction\s+\w+\s*\(", ln) and "end function" not in ln:
                stack.append("function")
        assert not stack, f"{fn}: unclosed blocks {stack}"
        for name in re.findall(r"bind\s*\(\s*c\s*,\s*name\s*=\s*'([^']+)'", txt):
            assert name in _lib.SYMBOLS, f"{fn}: bind(c) name {name} is not in the header"
