"""CPU tier for the PRUNED SEARCH of the roofline kernel: liliom_b200/csrc/knn_core.cuh compiled for the host (tests/knncore_host.cpp)
and compared, key for key, with an exhaustive search over the whole map.  What must hold: the running threshold,
the cell lower bounds, the trimming of runs, the per-thread run list and a tightened start threshold (temporal coherence) never
change the five smallest (fp32 distance, index) keys inside the gate; a query with fewer than five points inside the gate keeps
exactly those; points outside the 3x3x3 cell block are never inside the gate."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "build", "libknncore_host.so")
F = np.float32
EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)


@pytest.fixture(scope="module")
def kc():
    src = os.path.join(ROOT, "tests", "knncore_host.cpp")
    hdrs = [os.path.join(ROOT, "liliom_b200", "csrc", h) for h in ("knn_core.cuh", "dev_math.cuh", "ctx.cuh")]
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(p) for p in [src] + hdrs):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([gxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-D_GNU_SOURCE", "-Wno-attributes", "-Wno-unknown-pragmas",
                        "-I", cuda_inc, "-shared", "-o", SO, src], check=True)
    L = C.CDLL(SO)
    L.kc_gate_tau.argtypes = [C.c_double]; L.kc_gate_tau.restype = C.c_float
    L.kc_thread_knn5.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_float,
                                 C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_float, C.POINTER(C.c_uint64)]
    L.kc_thread_knn5.restype = C.c_uint64
    L.kc_owner_of.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_float]; L.kc_owner_of.restype = C.c_int
    return L


def _grid(pts, cell=1.0):
    """The cell grid of grid_build (liliom_b200/csrc/grid_knn.cu): cells floor(p / cell), x fastest, dense cell_start."""
    inv = F(1.0 / cell)
    c = np.floor(pts * inv).astype(np.int64)
    org = c.min(0); dim = c.max(0) - org + 1
    rel = c - org
    key = (rel[:, 2] * dim[1] + rel[:, 1]) * dim[0] + rel[:, 0]
    order = np.argsort(key, kind="stable")
    ms = np.zeros((len(pts), 4), F)
    ms[:, :3] = pts[order]
    ms[:, 3] = order.astype(np.int32).view(F)                      # w = index into the un-sorted map
    ncells = int(dim.prod())
    cell_start = np.searchsorted(key[order], np.arange(ncells + 1)).astype(np.int32)
    return ms, cell_start, org.astype(np.int32), dim.astype(np.int32), inv


def _exhaustive(q, pts, tau):
    """Five smallest (bits(d) << 32 | index) keys with d <= tau; d = ((dx*dx) + dy*dy) + dz*dz in fp32, dx = q - p (FLANN L2_Simple)."""
    d = pts.astype(F)
    dx = F(q[0]) - d[:, 0]; dy = F(q[1]) - d[:, 1]; dz = F(q[2]) - d[:, 2]
    dist = (dx * dx + dy * dy) + dz * dz
    assert dist.dtype == F
    keep = np.nonzero(dist <= tau)[0]
    keys = (dist[keep].view(np.uint32).astype(np.uint64) << np.uint64(32)) | keep.astype(np.uint64)
    keys.sort()
    out = np.full(5, EMPTY, np.uint64)
    out[: min(5, len(keys))] = keys[:5]
    return out, dist


def _search(kc, q, grid, tau):
    ms, cs, org, dim, inv = grid
    out = np.zeros(5, np.uint64)
    cand = kc.kc_thread_knn5(F(q[0]), F(q[1]), F(q[2]), ms.ctypes.data_as(C.POINTER(C.c_float)), cs.ctypes.data_as(C.POINTER(C.c_int)), inv,
                             org.ctypes.data_as(C.POINTER(C.c_int)), dim.ctypes.data_as(C.POINTER(C.c_int)), F(tau),
                             out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out, int(cand)


@pytest.mark.parametrize("seed,density", [(0, 3.0), (1, 12.0), (2, 40.0), (3, 0.6)])
def test_pruned_search_equals_exhaustive_search(kc, seed, density):
    rng = np.random.default_rng(seed)
    ext = np.array([9.0, 7.0, 4.0])
    m = int(density * ext.prod())
    pts = (rng.uniform(-0.5, 0.5, (m, 3)) * ext + np.array([100.3, -40.7, 2.2])).astype(F)
    # adversarial structure: exact duplicates (index tie-break), points on cell faces, a lattice with equal distances
    pts[: m // 20] = pts[m // 20: 2 * (m // 20)]
    pts[-(m // 25):] = np.round(pts[-(m // 25):])
    lat = np.stack(np.meshgrid(np.arange(-2, 3), np.arange(-2, 3), np.arange(-1, 2), indexing="ij"), -1).reshape(-1, 3) * 0.25
    pts = np.concatenate([pts, (lat + np.array([100.0, -41.0, 2.0])).astype(F)])
    grid = _grid(pts)
    gate = kc.kc_gate_tau(1.0)
    assert gate < 1.0 and np.nextafter(F(gate), F(2.0)) >= 1.0
    queries = np.concatenate([
        (rng.uniform(-0.55, 0.55, (400, 3)) * ext + np.array([100.3, -40.7, 2.2])),     # some outside the grid
        np.round(rng.uniform(-0.5, 0.5, (60, 3)) * ext + np.array([100.3, -40.7, 2.2])),  # on cell corners
        lat[:40] + np.array([100.0, -41.0, 2.0]) + 0.125,                                  # equidistant from lattice points
    ]).astype(F)
    examined = []
    n_full = 0
    for q in queries:
        want, dist = _exhaustive(q, pts, gate)
        got, cand = _search(kc, q, grid, gate)
        assert np.array_equal(got, want), (q, got, want)
        examined.append(cand)
        if want[4] != EMPTY:
            n_full += 1
            # temporal coherence: any start threshold >= the true fifth distance gives the same five keys and examines no more
            d5 = F(np.uint32(want[4] >> np.uint64(32)).view(F))
            for t in (d5, np.nextafter(d5, F(2.0)), F(min(float(d5) * 1.5 + 1e-3, float(gate)))):
                got2, cand2 = _search(kc, q, grid, t)
                assert np.array_equal(got2, want), (q, t)
                assert cand2 <= cand
        # exactness of the 3x3x3 block: nothing outside it is inside the gate
        c = np.floor(pts).astype(np.int64); cq = np.floor(q).astype(np.int64)
        outside = (np.abs(c - cq) > 1).any(1)
        assert not (dist[outside] <= gate).any()
    assert n_full > 50 or density < 1.0
    # the pruning prunes: on dense maps most of the block is never fetched
    if density >= 12.0:
        block = np.array([((np.abs(np.floor(pts).astype(np.int64) - np.floor(q).astype(np.int64)) <= 1).all(1)).sum() for q in queries[:100]])
        assert np.mean(examined[:100]) < 0.8 * np.mean(block)


def test_partition_rule_mirror_equals_the_device_function(kc):
    """liliom_b200/sharding.py::owner_of (what the gloo tests and the bench's shard report use) against knn_core.cuh::owner_of compiled
    for the host: same rank for points on both sides of cube faces, negative coordinates, the half-cube z shift, every rank count."""
    from liliom_b200 import sharding
    rng = np.random.default_rng(11)
    pts = np.concatenate([
        rng.uniform(-700, 700, (3000, 3)),
        np.round(rng.uniform(-40, 40, (600, 3))) * 16.0 + rng.choice([-1e-3, 0.0, 1e-3], (600, 3)),          # on / beside 16 m faces
        np.round(rng.uniform(-10, 10, (600, 3))) * 64.0 + np.array([0.0, 0.0, 32.0]) + rng.choice([-1e-3, 0.0, 1e-3], (600, 3)),
    ]).astype(F)
    for block in (16, 64):
        for nranks in (1, 2, 3, 4, 8, 16):
            want = sharding.owner_of(pts, nranks, block)
            got = np.array([kc.kc_owner_of(F(p[0]), F(p[1]), F(p[2]), nranks, F(1.0 / block)) for p in pts], np.int32)
            assert np.array_equal(got, want), (block, nranks)
            assert got.min() >= 0 and got.max() < nranks
            if nranks > 1:
                share = np.bincount(got[:3000], minlength=nranks) / 3000.0
                assert share.max() < 2.2 / nranks                                   # the linear hash spreads a 1.4 km stretch evenly
