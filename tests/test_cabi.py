"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol include/real_b200.h declares,
agrees on struct layouts, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest

from real_robots_b200 import _lib as L

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def header_functions():
    txt = open(os.path.join(ROOT, "include", "real_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z_0-9]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return L.load()


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libreal_b200.so does not export %s" % n
    # the Python binding covers the same set
    assert set(names) == set(L.SYMBOLS.keys())


def test_config_layout_and_defaults(lib):
    assert lib.rr_config_size() == C.sizeof(L.RRConfig)
    cfg = L.default_config()
    assert cfg.dt == 0.005 and cfg.solver_iterations == 50 and cfg.gravity == -9.81
    assert cfg.motor_kp == 0.1 and cfg.motor_kd == 1.0 and cfg.contact_breaking == 0.02


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    rc = lib.rr_create(C.byref(ctx), 0, 4, 3, L.DEFAULT_BLOB.encode(), None)
    assert rc != 0 and not ctx.value
    assert b"no CUDA device" in lib.rr_last_error(None)
    from real_robots_b200.sim import BatchedSim
    with pytest.raises(L.NativeLibraryError):
        BatchedSim(4)


def test_bad_arguments_are_rejected(lib):
    ctx = C.c_void_p()
    assert lib.rr_create(C.byref(ctx), 0, 0, 3, L.DEFAULT_BLOB.encode(), None) != 0
    assert lib.rr_create(C.byref(ctx), 0, 4, 7, L.DEFAULT_BLOB.encode(), None) != 0


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/real_b200.h must be consumable by a C compiler (the boundary is a C ABI, no C++ / torch types), and a C
    program using every entry point must link against libreal_b200.so."""
    import re
    import subprocess
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    hdr = os.path.join(root, "include", "real_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    names = sorted(set(re.findall(r"\b(rr_[a-z_]+)\s*\(", open(hdr).read())))
    src = tmp_path / "link.c"
    src.write_text('#include "real_b200.h"\n#include <stdio.h>\nint main(void) {\n  void* p[] = {%s};\n'
                   '  printf("%%d\\n", (int)(sizeof p / sizeof p[0]));\n  return rr_config_size() > 0 ? 0 : 1;\n}\n'
                   % ", ".join("(void*)%s" % n for n in names))
    exe = tmp_path / "link"
    libdir = os.path.join(root, "real_robots_b200")
    subprocess.check_call(["gcc", "-std=gnu99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libreal_b200.so", "-Wl,-rpath," + libdir,
                           "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and int(out.stdout) == len(names) >= 25
