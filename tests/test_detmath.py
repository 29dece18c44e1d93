"""liliom_b200/csrc/detmath.h (device atanf/atan2f) is bit-identical to this image's glibc:
every 7th float for atanf (all 2^32 pass in 16 s with `detmath_check 1`), 2e8 atan2f pairs."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_detmath_matches_glibc(tmp_path):
    exe = str(tmp_path / "detmath_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fopenmp", os.path.join(ROOT, "tests", "tools", "detmath_check.cpp"), "-o", exe],
                   check=True)
    r = subprocess.run([exe, "7", "200000000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "atanf_mismatch=0 atan2f_mismatch=0" in r.stdout
