"""The C-ABI shared library: loads without a GPU, exports every symbol include/qwen3_cuda.h declares,
and its checkpoint validation (which runs before any CUDA call) reports the reference's errors."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from qwen3_rs_b200 import transformer as T

HEADER = os.path.join(ROOT, "include", "qwen3_cuda.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(q3_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_trait_surface():
    syms = declared_symbols()
    for must in ("q3_create", "q3_forward", "q3_get_config", "q3_destroy", "q3_last_error"):
        assert must in syms


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C99 (no C++ types in the signatures) and link from C."""
    import shutil
    import subprocess
    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "qwen3_cuda.h"\n#include <stdio.h>\n'
                   'int main(void) { q3_handle *h = 0; int rc = q3_create("/nonexistent.bin", 0, 0, &h);\n'
                   '  printf("%d %s\\n", rc, q3_version()); return rc == Q3_EIO ? 0 : 1; }\n')
    lib = T._build.build()
    exe = tmp_path / "abi"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", str(src), "-I", os.path.join(ROOT, "include"),
                           "-L", os.path.dirname(lib), "-lqwen3cuda", "-Wl,-rpath," + os.path.dirname(lib), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.split()[0] == "-2"


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(T._build.build())
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in qwen3_cuda.h but not exported"


def test_python_binding_covers_the_whole_abi():
    assert sorted(T.ABI) == declared_symbols()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "qwen3_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for pat in (r"^\s*(from|import)\s+oracle\b", r"libq3oracle", r"q3_oracle\.c", r"np_forward"):
                    assert not re.search(pat, text, flags=re.M), f"{f} reaches into oracle/ ({pat})"


def _create(path, ctx=0):
    lib = T.load_library()
    h = C.c_void_p(0)
    rc = lib.q3_create(str(path).encode(), ctx, 0, C.byref(h))
    return rc, lib.q3_last_error().decode(), h


def test_missing_file_is_eio(tmp_path):
    rc, msg, _ = _create(tmp_path / "nope.bin")
    assert rc == -2 and "Failed to open checkpoint" in msg  # models/mod.rs:56-57


def test_bad_header_is_eformat(tmp_path, ckpt):
    raw = bytearray(open(ckpt("micro", 32, 7), "rb").read())
    cases = [((0, b"\0\0\0\0"), "Invalid checkpoint magic number"),  # configuration.rs:117-120
             ((4, b"\3\0\0\0"), "Unsupported checkpoint version"),   # :122-125
             ((12, b"\xff\xff\xff\xff"), "Invalid dim: must be positive"),  # :139-143
             ((8, b"\2\0\0\0"), "Unknown architecture_id: 2")]      # models/mod.rs:71
    for (off, val), want in cases:
        r = bytearray(raw)
        r[off:off + 4] = val
        p = tmp_path / "bad.bin"
        p.write_bytes(bytes(r))
        rc, msg, _ = _create(p)
        assert rc == -3 and want in msg, (rc, msg)
    p = tmp_path / "short.bin"
    p.write_bytes(bytes(raw[:1000]))
    rc, msg, _ = _create(p)
    assert rc == -3 and "Insufficient data" in msg  # utils.rs:21-26
    p.write_bytes(bytes(raw[:20]))
    rc, msg, _ = _create(p)
    assert rc == -3


def test_unsupported_shapes_fail_loudly(tmp_path, ckpt):
    raw = bytearray(open(ckpt("micro", 32, 7), "rb").read())
    raw[40:44] = (64).to_bytes(4, "little")  # head_dim 64: kernels are built for 128
    p = tmp_path / "hd64.bin"
    p.write_bytes(bytes(raw))
    rc, msg, _ = _create(p)
    assert rc in (-3, -5)


@pytest.mark.skipif(__import__("conftest").has_cuda(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_error_not_fallback(ckpt):
    """The forward path has no CPU fallback: without a device construction fails with Q3_ECUDA."""
    with pytest.raises(T.Q3Error) as e:
        T.TransformerBuilder.new(ckpt("micro", 32, 7)).build()
    assert e.value.code == -4
