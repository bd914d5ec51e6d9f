"""The C-ABI shared library loads on a machine without a GPU, exports every symbol that
include/b200snark.h declares, and refuses to compute without an sm_100 device (no CPU fallback)."""
import os
import re

import pytest

import snark_b200
from snark_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "b200snark.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = snark_b200.load_library()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200snark.h but not exported"
    assert sorted(L.SIGNATURES) == names, "snark_b200/lib.py binds exactly the header's functions"
    assert b"sm_100a" in lib.b2s_version()


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(snark_b200.B2SError) as e:
        snark_b200.Backend()
    assert e.value.code == 17  # B2S_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "snark_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle", src, flags=re.M), (dirpath, f)


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/b200snark.h must compile as C99 and link against the library from C."""
    import subprocess

    src = tmp_path / "use.c"
    src.write_text('#include "b200snark.h"\n#include <stdio.h>\n'
                   "int main(void) { b2s_ctx* c = 0; int32_t st = b2s_ctx_create(B2S_CURVE_BN254, 0, &c);\n"
                   '  printf("%s %d\\n", b2s_version(), (int)st); if (c) b2s_ctx_destroy(c); return 0; }\n')
    exe = tmp_path / "use"
    libdir = os.path.join(ROOT, "snark_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lb200snark", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "b200snark" in out.stdout
