"""CPU-only: the C++ facade that a maintainer adds to the reference build compiles against the
reference's UNMODIFIED APD.h / main.h (with the OpenCV/Boost header shim; neither library is installed)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "APD.h")), reason="reference tree not present on this box")
def test_facade_compiles_against_reference_header():
    gxx = shutil.which("g++")
    assert gxx
    cmd = [gxx, "-std=c++17", "-fsyntax-only", f"-I{ROOT}/oracle/shim", f"-I{REF}", f"-I{ROOT}/include",
           "-I/usr/local/cuda/include", f"{ROOT}/apd_mvs_b200/facade/APD_b200.cpp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_header_is_plain_c():
    gcc = shutil.which("gcc")
    src = '#include "apd_b200.h"\nint main(void){ apd_params p; apd_camera c; (void)p; (void)c; return sizeof(apd_params)==72 && sizeof(apd_camera)==112 ? 0 : 1; }\n'
    exe = "/tmp/apd_hdr_check"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", f"-I{ROOT}/include", "-x", "c", "-", "-o", exe], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([exe]).returncode == 0
