"""Device primitives (scan, stable radix sort, exact hash group-by) against numpy."""
import ctypes as C
import numpy as np
import pytest
from arriba_b200 import lib as L


def _fns(path):
    lib = L.load(path)
    p32 = C.POINTER(C.c_uint32)
    lib.arb_selftest_scan.argtypes = [p32, p32, C.c_uint32]
    lib.arb_selftest_sort.argtypes = [p32, p32, C.c_uint32, C.c_uint32]
    lib.arb_selftest_group.argtypes = [p32, p32, C.c_uint32]
    return lib


def check_prims(path, sizes):
    lib = _fns(path)
    rng = np.random.default_rng(5)
    for n in sizes:
        a = rng.integers(0, 1000, n, dtype=np.uint32)
        out = np.zeros(n + 1, np.uint32)
        assert lib.arb_selftest_scan(L.ptr(a), L.ptr(out), n) == 0
        want = np.concatenate([[0], np.cumsum(a, dtype=np.uint64)]).astype(np.uint32)
        assert np.array_equal(out, want), "scan n=%d" % n
        for bits in (1, 8, 13, 24, 32):
            hi = (1 << bits) - 1
            k = rng.integers(0, hi + 1, n, dtype=np.uint64).astype(np.uint32)
            if n > 10:
                k[: n // 3] = k[0]  # long runs of equal keys: stability matters
            v = np.arange(n, dtype=np.uint32)
            k2, v2 = k.copy(), v.copy()
            assert lib.arb_selftest_sort(L.ptr(k2), L.ptr(v2), n, bits) == 0
            order = np.argsort(k, kind="stable")
            assert np.array_equal(k2, k[order]) and np.array_equal(v2, v[order]), "sort n=%d bits=%d" % (n, bits)
        g = rng.integers(0, max(1, n // 7), n, dtype=np.uint32)
        first = np.zeros(max(n, 1), np.uint32)
        assert lib.arb_selftest_group(L.ptr(g), L.ptr(first), n) == 0
        if n:
            _, idx, inv = np.unique(g, return_index=True, return_inverse=True)
            assert np.array_equal(first[:n], idx[inv].astype(np.uint32)), "group n=%d" % n


def test_prims_hostsim(hostsim_lib):
    check_prims(hostsim_lib, [0, 1, 2, 33, 2047, 2048, 2049, 10000])


@pytest.mark.gpu
def test_prims_cuda(cuda_lib):
    check_prims(cuda_lib, [0, 1, 2, 33, 255, 256, 2047, 2048, 2049, 8192, 8193, 100000, 3000001])


def test_pool_size_classes(hostsim_lib):
    """Blocks are carved back to back out of slabs: every size class must keep the next block 512-byte aligned (a 20,000,000-byte class once did not)."""
    import ctypes as C
    lib = L.load(hostsim_lib)
    lib.arb_selftest_pool_size_class.argtypes = [C.c_uint64]; lib.arb_selftest_pool_size_class.restype = C.c_uint64
    rng = np.random.default_rng(3)
    sizes = np.unique(np.concatenate([np.arange(1, 5000, 37), (10 ** rng.uniform(3, 10.5, 4000)).astype(np.int64), 2 ** np.arange(9, 35)]))
    prev = 0
    for s in sizes:
        c = int(lib.arb_selftest_pool_size_class(int(s)))
        assert c >= s and c % 512 == 0 and c <= max(512, int(s) * 2), (s, c)
        assert c >= prev; prev = c


def test_annotation_query_front_end(hostsim_lib):
    """index_query.h asks with a 48-entry set first and repeats the query with the large set where that overflows: same answers as the plain query, for
    loci below, at and far above the small capacity."""
    lib = C.CDLL(hostsim_lib)
    lib.arb_selftest_index_query.argtypes = [C.c_uint32]
    for crowd in (3, 24, 47, 48, 49, 96, 1000):
        assert lib.arb_selftest_index_query(crowd) == 0, crowd


def test_host_sort_split_merges(hostsim_lib):
    """The host sort of the event stages (slice sorts, then merges cut by output rank through a spare buffer) against std::sort: thread counts that are not
    powers of two, sizes around the serial threshold, many equal keys."""
    lib = C.CDLL(hostsim_lib)
    lib.arb_selftest_host_sort.argtypes = [C.c_uint32, C.c_int, C.c_uint32]
    for threads in (1, 2, 3, 5, 8, 13, 32):
        for n in (0, 1, 100, 8191, 8192, 8193, 50_000, 300_001):
            assert lib.arb_selftest_host_sort(n, threads, 7 + n) == 0, (threads, n)
