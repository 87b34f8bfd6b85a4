"""The hand-written onesweep radix sort against numpy's stable sort (bit-exact permutation)."""
import ctypes as C

import numpy as np
import pytest
import torch

from ocrfdet_b200 import _lib

pytestmark = pytest.mark.gpu


def _sort(keys, vals, end_bit):
    L = _lib.lib()
    n = keys.shape[0]
    k = torch.from_numpy(keys.view(np.int64)).cuda()
    v = torch.from_numpy(vals.view(np.int32)).cuda()
    ko, vo, kt, vt = torch.empty_like(k), torch.empty_like(v), torch.empty_like(k), torch.empty_like(v)
    ws = torch.empty(int(L.ocrf_sort_workspace_bytes(max(n, 1))) + 256, dtype=torch.uint8, device="cuda")
    _lib.check(L.ocrf_sort_pairs(_lib.current_stream(), n, end_bit, _lib.ptr(k), _lib.ptr(v), _lib.ptr(ko), _lib.ptr(vo),
                                 _lib.ptr(kt), _lib.ptr(vt), _lib.ptr(ws)), "ocrf_sort_pairs")
    torch.cuda.synchronize()
    return ko.cpu().numpy().view(np.uint64), vo.cpu().numpy().view(np.uint32)


def _expect(keys, vals, end_bit):
    mask = np.uint64((1 << end_bit) - 1) if end_bit < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    order = np.argsort(keys & mask, kind="stable")
    return keys[order], vals[order]


@pytest.mark.parametrize("n", [1, 2, 31, 1023, 1025, 4095, 4096, 4097, 50_000, 1_000_003, 1_048_576, 1_048_577, 2_100_000])
@pytest.mark.parametrize("end_bit", [42, 44, 48, 64, 13])
def test_sort_pairs_matches_stable_numpy(n, end_bit):
    rng = np.random.default_rng(n * 131 + end_bit)
    # tile-like upper word, depth-like lower word with MANY exact ties (stability matters)
    tile = rng.integers(0, 2816, size=n, dtype=np.uint64)
    depth = rng.choice(rng.uniform(0.2, 80.0, size=max(4, n // 50)).astype(np.float32), size=n).view(np.uint32)
    keys = (tile << np.uint64(32)) | depth.astype(np.uint64)
    if end_bit == 64:
        keys |= rng.integers(0, 1 << 20, size=n, dtype=np.uint64) << np.uint64(44)
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = _sort(keys, vals, end_bit)
    wk, wv = _expect(keys, vals, end_bit)
    assert np.array_equal(gv, wv), "permutation differs (stability / order)"
    assert np.array_equal(gk, wk)


def test_sort_already_sorted_and_constant_keys():
    n = 20_000
    vals = np.arange(n, dtype=np.uint32)
    for keys in (np.arange(n, dtype=np.uint64) << np.uint64(20), np.full(n, 0x1234_0000_5678, dtype=np.uint64),
                 (np.arange(n, dtype=np.uint64)[::-1].copy() << np.uint64(16))):
        gk, gv = _sort(keys, vals, 48)
        wk, wv = _expect(keys, vals, 48)
        assert np.array_equal(gk, wk) and np.array_equal(gv, wv)


def test_sort_empty_is_a_noop():
    L = _lib.lib()
    assert L.ocrf_sort_pairs(_lib.current_stream(), 0, 42, None, None, None, None, None, None, None) == 0
