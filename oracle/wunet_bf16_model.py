"""
oracle/wunet_bf16_model.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The arithmetic of the bf16 tensor-core path (csrc/wunet_tc.cu) restated on the CPU, rounding where the kernels round:

* first encoder block: fp32 operands (CUDA cores), folded BatchNorm scale/shift in fp32, LeakyReLU, bf16 store; with
  ``enc0_tc`` (the library's tensor-core form of that block, 3 taps over groups of 8 samples) bf16-rounded samples and weights
  like every other block;
* every other block: bf16 activations x bf16 weights (round-to-nearest-even of the fp32 parameters), wide accumulation
  (float64 here; the tensor core accumulates exact products in fp32 — the difference is ~1e-6 relative and only matters
  where it flips a bf16 rounding), fp32 scale/shift + LeakyReLU, bf16 store;
* decoder inputs: the interpolated half in packed-bf16 arithmetic (lam = bf16(fma(up_scale, l, -m)), d = bf16(b - a),
  r = bf16(fma(lam, d, a))) for levels of at least 128 samples, in fp32 (lam0*f0 + lam1*f1, then bf16) for the packed
  short levels — exactly the two producer code paths;
* decoder blocks of at most 16 samples run as one dense GEMM over frames (gemm_tc_kernel) whose weights for the upsampled input
  have the interpolation folded in: W_eff = sum_t w[t] U[l+t-2, m] in fp32, rounded to bf16, applied to the previous block's
  bf16 rows directly (no rounding of interpolated values);
* head: the last block's fp32 (unrounded) activations, fp32 fma chain, tanh.

A GPU result that differs from this model by more than a bf16 ulp at a handful of rounding flips is a kernel bug; the
fp32 oracle alone cannot show that (the bf16 path sits 3e-3 of the level range away from it).  Only ``tests/`` may
import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .wunet_oracle import BN_EPS, LRELU_SLOPE, channel_plan


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _fold(st, prefix):
    """scale / shift exactly as pack_fp32_kernel computes them (fp32)."""
    g = torch.from_numpy(np.asarray(st[f"{prefix}.1.weight"], np.float32))
    beta = torch.from_numpy(np.asarray(st[f"{prefix}.1.bias"], np.float32))
    mean = torch.from_numpy(np.asarray(st[f"{prefix}.1.running_mean"], np.float32))
    var = torch.from_numpy(np.asarray(st[f"{prefix}.1.running_var"], np.float32))
    bias = torch.from_numpy(np.asarray(st[f"{prefix}.0.bias"], np.float32))
    s = g / torch.sqrt(var + np.float32(BN_EPS))
    shift = torch.addcmul(beta, bias - mean, s)               # fmaf(bias - mean, s, beta) up to one rounding
    return s, shift


def _block(x: torch.Tensor, st, prefix: str, k: int, quantise_operands: bool) -> torch.Tensor:
    """conv (wide accumulate) -> fp32 scale/shift -> LeakyReLU, fp32 result (not yet rounded to bf16)."""
    w = torch.from_numpy(np.asarray(st[f"{prefix}.0.weight"], np.float32))
    if quantise_operands:
        w = _bf16(w)
    acc = F.conv1d(x.double(), w.double(), None, padding=(k - 1) // 2).float()
    s, shift = _fold(st, prefix)
    v = acc * s[None, :, None] + shift[None, :, None]
    return torch.where(v >= 0, v, np.float32(LRELU_SLOPE) * v)


def _upsample_hfma2(prev: torch.Tensor) -> torch.Tensor:
    """Producer fast path (levels >= 128 samples): packed-bf16 interpolation between prev rows (l-1)>>1 and that + 1."""
    Lin = prev.shape[2]
    L = 2 * Lin
    up_scale = np.float32(Lin - 1) / np.float32(L - 1)
    l = torch.arange(L)
    m0 = torch.div(l - 1, 2, rounding_mode="floor")
    a = prev[:, :, m0.clamp(0, Lin - 1)]
    b = prev[:, :, (m0 + 1).clamp(0, Lin - 1)]
    lam = _bf16((l.double() * float(up_scale) - m0.double()).float())       # one fused rounding to fp32, then bf16
    d = _bf16(b - a)
    return _bf16((lam[None, None, :].double() * d.double() + a.double()).float())


def _block_folded(prev: torch.Tensor, skip: torch.Tensor, st, prefix: str) -> torch.Tensor:
    """Decoder block of <= 16 samples as gemm_tc_kernel computes it: interpolation folded into the weights of the upsampled
    input (expand_gemm_weights_kernel), skip half as a plain 5-tap convolution; fp32 result before the bf16 store."""
    w = torch.from_numpy(np.asarray(st[f"{prefix}.0.weight"], np.float32))            # [Cout][Cprev + Cskip][5]
    Cprev, Lin = prev.shape[1], prev.shape[2]
    L = 2 * Lin
    up_scale = np.float32(Lin - 1) / np.float32(L - 1) if L > 1 else np.float32(0)
    sidx = (torch.arange(L, dtype=torch.float32) * up_scale)
    i0 = sidx.to(torch.int64)
    i1 = i0 + (i0 < Lin - 1).to(torch.int64)
    lam1 = (sidx - i0.to(torch.float32)).double()
    U = torch.zeros(L, Lin, dtype=torch.float64)
    U[torch.arange(L), i0] += 1.0 - lam1
    U[torch.arange(L), i1] += lam1
    wp = w[:, :Cprev, :].double()
    weff = torch.zeros(w.shape[0], Cprev, L, Lin, dtype=torch.float64)                 # [co][ci][l][m]
    for t in range(5):
        for l in range(L):
            lp = l + t - 2
            if 0 <= lp < L:
                weff[:, :, l, :] += wp[:, :, t, None] * U[lp][None, None, :]
    weff = _bf16(weff.float()).double()
    acc = torch.einsum("bim,oilm->bol", prev.double(), weff)
    acc = acc + F.conv1d(skip.double(), _bf16(w[:, Cprev:, :]).double(), None, padding=2)
    s, shift = _fold(st, prefix)
    v = acc.float() * s[None, :, None] + shift[None, :, None]
    return torch.where(v >= 0, v, np.float32(LRELU_SLOPE) * v)


def _upsample_fp32(prev: torch.Tensor) -> torch.Tensor:
    """Producer generic path (packed short levels): ATen index math and lam0*f0 + lam1*f1 in fp32, then bf16."""
    Lin = prev.shape[2]
    L = 2 * Lin
    up_scale = np.float32(Lin - 1) / np.float32(L - 1) if L > 1 else np.float32(0)
    s = (torch.arange(L, dtype=torch.float32) * up_scale)
    i0 = s.to(torch.int64)
    i1 = i0 + (i0 < Lin - 1).to(torch.int64)
    lam1 = s - i0.to(torch.float32)
    lam0 = 1.0 - lam1
    return _bf16(lam0[None, None, :] * prev[:, :, i0] + lam1[None, None, :] * prev[:, :, i1])


def forward_bf16_model(state, x: np.ndarray, n_layers: int = 12, channels_interval: int = 24, return_levels: bool = False,
                       forced=None, dense_bottom=None, enc0_tc: bool = False):
    """-> y [B,1,T] fp32 (and the bf16-valued outputs of the 2n+1 blocks as fp32 arrays; the last one unrounded).

    ``dense_bottom``: whether decoder blocks of at most 16 samples use the folded-interpolation weights of gemm_tc_kernel (the
    library takes that kernel for batches of at least 64 frames; default: decided from the batch size like the library).

    ``forced``: optional list of the 2n block outputs measured on the GPU (fp32 arrays holding bf16 values). Block i is then
    evaluated on the GPU's outputs of the blocks before it ("teacher forcing"): rounding flips do not propagate, so every
    level can be held to bit-equality up to its own flips, however deep it is."""
    plan = channel_plan(n_layers, channels_interval)
    n = n_layers
    xt = torch.from_numpy(np.asarray(x, np.float32))
    if dense_bottom is None:
        dense_bottom = xt.shape[0] >= 64
    levels, skips = [], []

    def keep(i, t):
        levels.append(t)
        return torch.from_numpy(np.asarray(forced[i], np.float32)) if (forced is not None and i < 2 * n) else t

    o = keep(0, _bf16(_block(_bf16(xt) if enc0_tc else xt, state, plan[0][0], 15, quantise_operands=enc0_tc)))
    skips.append(o)
    o = o[:, :, ::2]
    for i in range(1, n):
        o = keep(i, _bf16(_block(o, state, plan[i][0], 15, True)))
        skips.append(o)
        o = o[:, :, ::2]
    o = keep(n, _bf16(_block(o, state, "middle", 15, True)))
    for j in range(n):
        L = 2 * o.shape[2]
        if dense_bottom and L <= 16 and j != n - 1:
            v = _block_folded(o, skips[n - 1 - j], state, plan[n + 1 + j][0])
        else:
            up = _upsample_hfma2(o) if L >= 128 else _upsample_fp32(o)
            v = _block(torch.cat([up, skips[n - 1 - j]], dim=1), state, plan[n + 1 + j][0], 5, True)
        o = v if j == n - 1 else keep(n + 1 + j, _bf16(v))      # the last block feeds the fused head unrounded
        if j == n - 1:
            levels.append(o)
    w = torch.from_numpy(np.asarray(state["out.0.weight"], np.float32))[0, :, 0]
    b = np.float32(np.asarray(state["out.0.bias"], np.float32)[0])
    C = o.shape[1]
    pre = (o.double() * w[:C].double()[None, :, None]).sum(dim=1) + w[C].double() * xt[:, 0].double() + float(b)
    y = torch.tanh(pre).float()[:, None, :].numpy()
    return (y, [t.numpy() for t in levels]) if return_levels else y
