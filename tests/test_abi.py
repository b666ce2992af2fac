"""CPU tests of the drop-in boundary: the shared library loads and exports every symbol
include/mhb_spgemm.h declares; without a GPU it refuses to run (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import mh_spgemm_b200  # noqa: F401
from mh_spgemm_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mhb_spgemm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mhb_[A-Za-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from importlib import import_module
    import_module("mh_spgemm_b200.build").build()
    L = C.CDLL(api.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mhb_spgemm.h but not exported"
    assert sorted(api.ABI_SYMBOLS) == names, "api.ABI_SYMBOLS out of sync with the header"


def test_version_string():
    L = api.load_library()
    assert b"sm_100a" in L.mhb_version()


def test_no_cpu_fallback_without_device():
    """On a box without a CUDA device the handle cannot be created; nothing silently
    routes to the oracle."""
    import torch
    if torch.cuda.is_available():
        return
    L = api.load_library()
    h = C.c_void_p()
    assert L.mhb_create(C.byref(h), 0) != 0
    assert not h.value
    src = open(os.path.join(ROOT, "mh-spgemm_b200", "api.py")).read()
    assert "import oracle" not in src and "from oracle" not in src


def test_compat_header_mentions_reference_interface():
    """The C++ shim keeps the reference's class and entry-point names (inc/CSR.h, src/main.cu:12)."""
    p = os.path.join(ROOT, "include", "mhb_compat.hpp")
    if not os.path.exists(p):
        return
    s = open(p).read()
    for token in ("class CSR", "class Tool", "class Timing", "MH_spgemm", "d_ptr", "d_col", "d_val", "H2D", "D2H"):
        assert token in s


def test_bin_ladders_are_consistent(tmp_path):
    """csrc/mhb_config.h compiled on the host: every (span, nnz, products, forced path) lands in a
    bin whose kernel can hold the row (hash fill <= 5/8 numeric, <= 3/4 symbolic; windows and
    bitmaps within the shared memory their launch requests)."""
    exe = tmp_path / "ladders"
    src = os.path.join(ROOT, "tests", "cpp", "ladders.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", str(exe), src])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the shipped package (Python or CUDA sources)
    may import, include or load it."""
    pkg = os.path.join(ROOT, "mh-spgemm_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                continue
            text = open(os.path.join(dirpath, f), encoding="utf-8", errors="replace").read()
            for needle in ("import oracle", "from oracle", "liboracle", "libmhref", "oracle/"):
                if needle in text:
                    bad.append((f, needle))
    assert not bad, bad


def test_header_is_plain_c(tmp_path):
    """include/mhb_spgemm.h is the FFI surface: it must compile as C99 with no C++ and no CUDA headers."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "mhb_spgemm.h"\n'
                   "int main(void) { mhb_handle_t h = 0; mhb_shard_t s = 0; mhb_timing t; mhb_stats st; (void)h; (void)s;\n"
                   "  (void)t; (void)st; return (MHB_OK == 0 && MHB_ERR_CAPACITY == 5 && MHB_SHARD_BLOB_BYTES == 128) ? 0 : 1; }\n")
    out = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                          "-o", str(tmp_path / "hdr"), str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert subprocess.run([str(tmp_path / "hdr")]).returncode == 0


def test_cpp_shard_driver_links():
    """The plain C++ caller of the row-sharded API (tests/cpp/shard_driver.cpp: g++ only, no CUDA headers)
    compiles against include/mhb_spgemm.h and links with the library; the -m gpu suite runs it."""
    from importlib import import_module
    exe = import_module("mh_spgemm_b200.build").build_shard_driver()
    assert os.path.exists(exe)
