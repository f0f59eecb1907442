"""The device StratifiedSampler (csrc/stratified.cuh, run on the host through rrt_stratified_host_probe) against the
oracle's PixelSampler<Stratified> (oracle/rt_sampling.hpp, orc_kat_stratified): the same PCG32 streams must give the
same tables and overflow draws bit for bit.  The distribution itself (strata, shuffle, U[-1,1) overflow) is checked
against stratified.rs in tests/test_reference_image.py::test_stratified_sampler_tables."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from rs_ray_toy_b200 import capi


@pytest.mark.parametrize("xs,ys,nd,jitter", [(4, 4, 4, 1), (4, 4, 4, 0), (3, 5, 2, 1), (1, 1, 1, 1), (16, 16, 3, 1), (2, 8, 7, 0)])
def test_device_sampler_equals_the_oracles(xs, ys, nd, jitter):
    Lo, Ld = O.lib(), capi.lib()
    Lo.orc_kat_stratified.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    Ld.rrt_stratified_host_probe.restype = C.c_int
    Ld.rrt_stratified_host_probe.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_uint32, C.c_uint32, C.c_uint32,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    n = xs * ys
    for seed, xres, px, py in [(1, 640, 0, 0), (7, 640, 17, 5), (123456789012345, 3840, 3839, 2159), (0, 1, 0, 0)]:
        ref = [np.zeros((nd, n)), np.zeros((nd, n, 2)), np.zeros((n, 4))]
        got = [np.zeros((nd, n)), np.zeros((nd, n, 2)), np.zeros((n, 4))]
        Lo.orc_kat_stratified(seed, xres, px, py, xs, ys, nd, jitter, *[a.ctypes.data for a in ref])
        capi.check(Ld.rrt_stratified_host_probe(seed, xres, px, py, xs, ys, nd, jitter, *[a.ctypes.data for a in got]))
        for a, b in zip(ref, got):
            assert np.array_equal(a, b)


def test_probe_refuses_tables_beyond_the_device_range():
    Ld = capi.lib()
    Ld.rrt_stratified_host_probe.restype = C.c_int
    Ld.rrt_stratified_host_probe.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_uint32, C.c_uint32, C.c_uint32,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    a = np.zeros(8)
    assert Ld.rrt_stratified_host_probe(1, 64, 0, 0, 32, 32, 1, 1, a.ctypes.data, a.ctypes.data, a.ctypes.data) == capi.RRT_ERR_UNSUPPORTED
