"""CPU-side checks of the drop-in boundary (no GPU): the C-ABI library loads, exports every symbol
include/liliom.h declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="liliom.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(liliom_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import liliom_b200 as L
    lib = L._binding.lib()
    decl = _declared_symbols()
    assert len(decl) >= 25
    missing = [s for s in decl if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(L.EXPORTS) == decl, (set(decl) ^ set(L.EXPORTS))
    nodes = _declared_symbols("liliom_nodes.h")
    assert sorted(L.NODE_EXPORTS) == nodes, (set(nodes) ^ set(L.NODE_EXPORTS))
    assert not [s for s in nodes if not hasattr(lib, s)]


def test_struct_layouts_match_header():
    import liliom_b200 as L
    assert L.PT48.itemsize == 48 and L.PT32.itemsize == 32
    assert ctypes.sizeof(L.IterStats) == 4 + 4 + 8 + 27 * 8 + 7 * 8
    assert ctypes.sizeof(L.Counters) == 48
    p = L.default_params(0)
    assert (p.abi_version, p.point_stride, p.surf_thres, p.edge_thres) == (1, 48, 0.2, 4.0)
    assert (p.knn_max_sqdist, p.plane_thres, p.weight_gate, p.huber_a, p.max_map_frames) == (1.0, 0.06, 0.4, 0.1, 20)
    assert abs(p.leaf_scan - 0.4) < 1e-7 and abs(p.leaf_map - 0.4) < 1e-7
    r = L.default_params(1)
    assert (r.point_stride, r.line_num, r.ds_rate) == (32, 64, 4) and abs(r.rot_ds_leaf - 0.6) < 1e-7


def test_strerror_and_bad_arguments():
    import liliom_b200 as L
    lib = L._binding.lib()
    assert lib.liliom_strerror(0) == b"ok"
    assert b"10" in lib.liliom_strerror(-3)            # "< 10" map points, LidarOdometry.cpp:485-488
    assert lib.liliom_create(None, None, 0) == -1
    h = ctypes.c_void_p()
    bad = L.default_params(0); bad.abi_version = 99
    assert lib.liliom_create(ctypes.byref(h), ctypes.byref(bad), 0) == -1
    bad = L.default_params(0); bad.point_stride = 40
    assert lib.liliom_create(ctypes.byref(h), ctypes.byref(bad), 0) == -1


def test_no_cpu_fallback_without_a_gpu():
    import torch
    import liliom_b200 as L
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(L.LiliomError) as e:
        L.Context()
    assert e.value.code == L._binding.E_CUDA


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under liliom_b200/ may import, load or link it."""
    for dp, _, files in os.walk(os.path.join(ROOT, "liliom_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)


def test_pointcloud2_layout_matches_the_point_dtypes():
    """liliom_pc2_layout (pcl::toROSMsg's field table) against the numpy mirrors of the PCL structs."""
    import liliom_b200 as L
    f48, step48 = L.pc2_layout(48)
    assert step48 == 48 and [x[0] for x in f48] == ["x", "y", "z", "normal_x", "normal_y", "normal_z", "intensity", "curvature"]
    names48 = {"normal_x": "nx", "normal_y": "ny", "normal_z": "nz"}
    for name, off, dt, cnt in f48:
        assert (dt, cnt) == (7, 1) and L.PT48.fields[names48.get(name, name)][1] == off
    f32, step32 = L.pc2_layout(32)
    assert step32 == 32 and [(n, o) for n, o, _, _ in f32] == [("x", 0), ("y", 4), ("z", 8), ("intensity", 16)]
    for name, off, dt, cnt in f32:
        assert L.PT32.fields[name][1] == off
    assert L._binding.lib().liliom_pc2_layout(40, None, 0, None) == L._binding.E_ARG
    assert L._binding.lib().liliom_pc2_layout(48, (L._binding.Pc2Field * 2)(), 2, None) == L._binding.E_CAPACITY


def test_headers_are_plain_c_and_link_from_c(tmp_path):
    """The drop-in boundary is a C ABI: both headers must compile as strict C99 and as C++14 (the reference's dialect,
    L/CMakeLists.txt:6), and a C program must link against the library and call it without any CUDA/torch type."""
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gcc or not gxx:
        pytest.skip("no host compiler")
    src = tmp_path / "hdr.c"
    src.write_text('#include "liliom.h"\n#include "liliom_nodes.h"\n'
                   'int main(void) { liliom_params p; liliom_pc2_field f[8]; int step = 0;\n'
                   '  liliom_default_params(&p, 1);\n'
                   '  if (p.abi_version != LILIOM_ABI_VERSION || p.point_stride != 32) return 1;\n'
                   '  if (liliom_pc2_layout(48, f, 8, &step) != 8 || step != 48) return 2;\n'
                   '  return liliom_strerror(LILIOM_E_FEWMAP)[0] ? 0 : 3; }\n')
    inc = os.path.join(ROOT, "include")
    libdir = os.path.join(ROOT, "liliom_b200")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, "-c", str(src), "-o", str(tmp_path / "a.o")], check=True)
    subprocess.run([gxx, "-std=c++14", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, "-x", "c++", "-c", str(src), "-o", str(tmp_path / "b.o")], check=True)
    exe = tmp_path / "hdr_bin"
    subprocess.run([gcc, "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-lliliom_b200", f"-Wl,-rpath,{libdir}"], check=True)
    assert subprocess.run([str(exe)]).returncode == 0
