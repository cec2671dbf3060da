"""Row-pair form of a conv block (WUNET_TC_PAIR): the weight expansion the library packs for the tensor
cores (wunet_debug_pair_weights, the host side of the device function pair_weight) must turn a K-tap Conv1d over positions
(model/unet_basic.py:10,23) into the SAME operator over pairs of positions. CPU only: the layouts are re-derived here in numpy
from their description in include/wunet_b200.h / wunet_tc.cu, independently of the C code."""
import ctypes

import numpy as np
import pytest

from wave_u_net_for_speech_enhancement_b200 import _lib


def conv1d_same(x, w):
    """x [L][Cin], w [Cout][Cin][K] -> y [L][Cout], zero padding (K-1)/2 like nn.Conv1d(padding=K//2)."""
    L, cin = x.shape
    cout, _, K = w.shape
    P = (K - 1) // 2
    xp = np.zeros((L + 2 * P, cin), dtype=np.float64)
    xp[P:P + L] = x
    y = np.zeros((L, cout), dtype=np.float64)
    for t in range(K):
        y += xp[t:t + L] @ w[:, :, t].T.astype(np.float64)
    return y


def pair_weights(w, c0, c1, dec):
    cout, cin, K = w.shape
    assert cin == c0 + c1
    P = (K - 1) // 2
    kp = 2 * ((P + 1) // 2) + 1
    out = np.zeros((2 * cout, 2 * cin, kp), dtype=np.float32)
    wc = np.ascontiguousarray(w, dtype=np.float32)
    _lib.check(_lib.load().wunet_debug_pair_weights(wc.ctypes.data_as(ctypes.c_void_p), cout, c0, c1, K, int(dec),
                                                     out.ctypes.data_as(ctypes.c_void_p)))
    return out, kp


def virtual_rows(x0, x1, dec):
    """Operand rows of the pair block: row m = positions 2m (q=0) and 2m+1 (q=1).
    Segment 1 and an encoder's segment 0: the contiguous view [L/2][2C] (v = q*C + c). A decoder's segment 0 (written by the
    producer warps): per 64-wide chunk k the real channels 32k .. 32k+w-1 (w = min(32, C0-32k)) as [q0 range | q1 range]."""
    L, c0 = x0.shape
    if not dec:
        v0 = x0.reshape(L // 2, 2 * c0)
    else:
        cols = []
        for lo in range(0, c0, 32):
            hi = min(c0, lo + 32)
            cols.append(x0[0::2, lo:hi])
            cols.append(x0[1::2, lo:hi])
        v0 = np.concatenate(cols, axis=1)
    assert v0.shape == (L // 2, 2 * c0)
    if x1 is None:
        return v0
    return np.concatenate([v0, x1.reshape(L // 2, 2 * x1.shape[1])], axis=1)


@pytest.mark.parametrize("c0,c1,cout,K,dec", [(24, 0, 48, 15, 0), (48, 24, 24, 5, 1), (8, 0, 16, 15, 0), (32, 8, 8, 5, 1),
                                               (72, 48, 48, 5, 1), (40, 16, 16, 5, 1)])
def test_pair_block_equals_conv(c0, c1, cout, K, dec):
    rng = np.random.default_rng(c0 * 131 + c1 * 17 + K)
    L = 64
    w = rng.standard_normal((cout, c0 + c1, K)).astype(np.float32)
    x0 = rng.standard_normal((L, c0)).astype(np.float32)
    x1 = rng.standard_normal((L, c1)).astype(np.float32) if c1 else None
    x = x0 if x1 is None else np.concatenate([x0, x1], axis=1)
    want = conv1d_same(x, w)                                            # [L][cout]
    wp, kp = pair_weights(w, c0, c1, dec)                               # [2cout][2cin][kp]
    assert kp == {15: 9, 5: 3}[K]
    xv = virtual_rows(x0, x1, dec)                                      # [L/2][2cin]
    got = conv1d_same(xv, wp)                                           # kp-tap conv over row pairs, zero rows outside the frame
    got = got.reshape(L // 2, 2, cout).reshape(L, cout)                 # column r*cout + co of row m = position 2m + r
    assert np.abs(got - want).max() < 1e-9 * max(1.0, np.abs(want).max())


def test_pair_weights_zero_pattern():
    """Taps that fall outside the original kernel are exactly zero (no wrap-around), and every original weight is used."""
    rng = np.random.default_rng(5)
    w = rng.standard_normal((4, 8, 5)).astype(np.float32) + 3.0          # no zeros in the source
    wp, kp = pair_weights(w, 8, 0, 0)
    # each original tap t of (co, c) appears once per output parity r (q = (t - P + r) mod 2 and dm follow): 2 copies in total
    assert np.count_nonzero(wp) == 2 * w.size
    assert sorted(np.unique(wp[wp != 0]).tolist()) == sorted(np.unique(w).tolist())


def test_group8_block_equals_conv():
    """Block 0, Conv1d(1 -> C, k=15), over groups of 8 samples (WUNET_TC_ENC0): 3 taps over rows of 8 samples, 8 C columns."""
    rng = np.random.default_rng(8)
    C, K, L = 24, 15, 256
    w = rng.standard_normal((C, 1, K)).astype(np.float32)
    x = rng.standard_normal((L, 1)).astype(np.float32)
    want = conv1d_same(x, w)                                            # [L][C]
    out = np.zeros((8 * C, 8, 3), dtype=np.float32)
    _lib.check(_lib.load().wunet_debug_pair_weights(w.ctypes.data_as(ctypes.c_void_p), C, 1, 0, K, 2, out.ctypes.data_as(ctypes.c_void_p)))
    got = conv1d_same(x.reshape(L // 8, 8), out)                        # [L/8][8 C]: column r*C + co of row m = sample 8m + r
    assert np.abs(got.reshape(L, C) - want).max() < 1e-9 * np.abs(want).max()
    assert np.count_nonzero(out) == 8 * w.size                          # every tap once per output phase r


def test_pair_block_equals_conv_property():
    """hypothesis: any channel counts (multiples of 8 as the library requires), both layouts, odd tap counts up to 15."""
    from hypothesis import given, settings, strategies as stt

    @settings(max_examples=40, deadline=None)
    @given(c0=stt.integers(1, 12).map(lambda v: 8 * v), c1=stt.integers(0, 6).map(lambda v: 8 * v), cout=stt.integers(1, 6).map(lambda v: 8 * v),
           K=stt.sampled_from([1, 3, 5, 7, 9, 15]), dec=stt.booleans(), L=stt.sampled_from([2, 8, 30]), seed=stt.integers(0, 2**16))
    def check(c0, c1, cout, K, dec, L, seed):
        if not dec:
            c1 = 0
        rng = np.random.default_rng(seed)
        w = rng.standard_normal((cout, c0 + c1, K)).astype(np.float32)
        x0 = rng.standard_normal((L, c0)).astype(np.float32)
        x1 = rng.standard_normal((L, c1)).astype(np.float32) if c1 else None
        x = x0 if x1 is None else np.concatenate([x0, x1], axis=1)
        wp, kp = pair_weights(w, c0, c1, dec)
        P = (K - 1) // 2
        assert kp == 2 * ((P + 1) // 2) + 1
        got = conv1d_same(virtual_rows(x0, x1, dec), wp).reshape(L // 2, 2, cout).reshape(L, cout)
        want = conv1d_same(x, w)
        assert np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max())

    check()
