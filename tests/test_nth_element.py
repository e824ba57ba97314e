"""The restated libstdc++ std::nth_element (eventcalib_b200/csrc/ecb_nth_element.h, used on the device to pick the
cluster "centre" exactly like CirclesEventFrame.cpp:140-147) against the real std::nth_element of this image's
libstdc++ — whole permuted arrays, heavy ties, sizes that reach the insertion-sort, partition and heap-select paths."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    so = os.path.join(ROOT, "tests", "_build", "libnth_host.so")
    src = os.path.join(ROOT, "tests", "helpers", "nth_element_host.cpp")
    hdr = os.path.join(ROOT, "eventcalib_b200", "csrc", "ecb_nth_element.h")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


def _both(lib, v, key, nth):
    v = np.ascontiguousarray(v, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    a = np.zeros_like(v)
    b = np.zeros_like(v)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    lib.nth_both(p(v), len(v), p(key), int(nth), p(a), p(b))
    return a, b


def test_matches_std_nth_element_with_ties():
    lib = _lib()
    rng = np.random.default_rng(0)
    for n in list(range(1, 40)) + [63, 64, 65, 100, 257, 1000, 5000]:
        for distinct in (1, 2, 3, 7, n // 2 + 1, 10 * n):
            for rep in range(6):
                key = rng.integers(0, distinct, n).astype(np.uint32)
                v = rng.permutation(n).astype(np.uint32)
                nth = n // 2 if rep < 4 else int(rng.integers(0, n))
                a, b = _both(lib, v, key, nth)
                assert np.array_equal(a, b), (n, distinct, rep)


def test_adversarial_patterns_reach_heap_select():
    """median-of-3 killers / organ pipes / sorted runs: some exhaust the 2*lg(n) depth limit (heap-select path)"""
    lib = _lib()
    for n in (64, 200, 1024, 4097):
        i = np.arange(n, dtype=np.uint32)
        pats = [i, i[::-1], np.minimum(i, n - 1 - i), np.maximum(i, n - 1 - i), (i * 7919) % 13, (i % 2) * n + i // 2,
                np.where(i % 2 == 0, i, n - i)]
        k = n // 2
        killer = np.zeros(n, np.uint32)      # Musser's median-of-3 killer sequence
        for j in range(k):
            killer[2 * j if j % 2 == 0 else 2 * j] = j + 1
        killer[1::2] = np.arange(k, k + len(killer[1::2]))
        pats.append(killer)
        for key in pats:
            key = np.ascontiguousarray(key, np.uint32)
            a, b = _both(lib, i, key, n // 2)
            assert np.array_equal(a, b)
