"""
oracle/wunet_train_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy, float64, written from the formulas — no autograd) of ONE training step of the reference
Wave-U-Net, the "next" row N1 of SURVEY.md §8f:

    trainer/trainer.py:34-38   enhanced = model(mixture); loss = loss_function(clean, enhanced); loss.backward()
    model/loss.py:3-4          nn.MSELoss()  (mean over all elements)
    model/unet_basic.py:77-100 forward, here with BatchNorm1d in TRAINING mode (:12, :25, :55 — batch statistics,
                               running-stat update with momentum 0.1 and the unbiased variance)

``forward_train`` returns the output, the updated running statistics and a cache; ``backward`` returns the gradient of
every parameter under the reference's ``state_dict`` key.  Pinned by ``tests/golden/train_*.npz``, which
``oracle/gen_golden_train.py`` produced with autograd on the *live reference module* (cast to float64).
Only ``tests/`` may import this module.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .wunet_oracle import BN_EPS, LRELU_SLOPE, channel_plan

BN_MOMENTUM = 0.1      # torch.nn.BatchNorm1d default


def _windows(xp: np.ndarray, K: int, L: int) -> np.ndarray:
    """[B, C, L + K - 1] -> view [B, C, L, K] of the K taps of every output position."""
    s = xp.strides
    return np.lib.stride_tricks.as_strided(xp, shape=(xp.shape[0], xp.shape[1], L, K), strides=(s[0], s[1], s[2], s[2]))


def conv1d_fwd(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    K = w.shape[2]
    pad = (K - 1) // 2
    xp = np.pad(x, ((0, 0), (0, 0), (pad, K - 1 - pad)))
    return np.einsum("bclk,ock->bol", _windows(xp, K, x.shape[2]), w, optimize=True) + b[None, :, None]


def conv1d_bwd(x: np.ndarray, w: np.ndarray, dz: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> (dx, dw, db) of z = conv1d(x, w, b) with 'same' zero padding."""
    K = w.shape[2]
    pad = (K - 1) // 2
    L = x.shape[2]
    xp = np.pad(x, ((0, 0), (0, 0), (pad, K - 1 - pad)))
    dw = np.einsum("bol,bclk->ock", dz, _windows(xp, K, L), optimize=True)
    db = dz.sum(axis=(0, 2))
    dxp = np.zeros_like(xp)
    for k in range(K):
        dxp[:, :, k:k + L] += np.einsum("bol,oc->bcl", dz, w[:, :, k], optimize=True)
    return dxp[:, :, pad:pad + L], dw, db


def upsample_matrix(Lin: int) -> np.ndarray:
    """U [2 Lin, Lin]: F.interpolate(scale_factor=2, mode='linear', align_corners=True) as a matrix (model/unet_basic.py:93)."""
    L = 2 * Lin
    U = np.zeros((L, Lin))
    if Lin == 1:
        U[:, 0] = 1.0
        return U
    src = np.arange(L) * ((Lin - 1) / (L - 1))
    i0 = np.minimum(np.floor(src).astype(int), Lin - 1)
    i1 = np.minimum(i0 + 1, Lin - 1)
    lam = src - i0
    U[np.arange(L), i0] += 1.0 - lam
    U[np.arange(L), i1] += lam
    return U


def _block_fwd(x, st, prefix, new_stats):
    """Conv1d -> BatchNorm1d (training) -> LeakyReLU; returns (activation, cache)."""
    w = st[f"{prefix}.0.weight"].astype(np.float64)
    b = st[f"{prefix}.0.bias"].astype(np.float64)
    g = st[f"{prefix}.1.weight"].astype(np.float64)
    be = st[f"{prefix}.1.bias"].astype(np.float64)
    z = conv1d_fwd(x, w, b)
    n = z.shape[0] * z.shape[2]
    mu = z.mean(axis=(0, 2))
    var = z.var(axis=(0, 2))                                   # biased: what normalises the batch
    inv = 1.0 / np.sqrt(var + BN_EPS)
    zh = (z - mu[None, :, None]) * inv[None, :, None]
    y = g[None, :, None] * zh + be[None, :, None]
    a = np.where(y >= 0, y, LRELU_SLOPE * y)
    rm = st[f"{prefix}.1.running_mean"].astype(np.float64)
    rv = st[f"{prefix}.1.running_var"].astype(np.float64)
    new_stats[f"{prefix}.1.running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mu
    new_stats[f"{prefix}.1.running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))
    return a, (x, w, g, zh, inv, y)


def _block_bwd(da, cache, prefix, grads):
    """Gradient of the block input; parameter gradients go into `grads`."""
    x, w, g, zh, inv, y = cache
    dy = np.where(y >= 0, da, LRELU_SLOPE * da)
    grads[f"{prefix}.1.bias"] = dy.sum(axis=(0, 2))
    grads[f"{prefix}.1.weight"] = (dy * zh).sum(axis=(0, 2))
    n = dy.shape[0] * dy.shape[2]
    dzh = dy * g[None, :, None]
    dz = inv[None, :, None] * (dzh - dzh.mean(axis=(0, 2))[None, :, None]
                               - zh * (dzh * zh).sum(axis=(0, 2))[None, :, None] / n)
    dx, dw, db = conv1d_bwd(x, w, dz)
    grads[f"{prefix}.0.weight"] = dw
    grads[f"{prefix}.0.bias"] = db
    return dx


def forward_train(state: Dict[str, np.ndarray], x: np.ndarray, n_layers: int = 12, channels_interval: int = 24):
    """-> (y [B,1,T], new running stats, cache). x: [B,1,T]."""
    plan = channel_plan(n_layers, channels_interval)
    n = n_layers
    x = x.astype(np.float64)
    new_stats: Dict[str, np.ndarray] = {}
    caches = []
    skips = []
    o = x
    for i in range(n):                                          # model/unet_basic.py:82-86
        o, c = _block_fwd(o, state, plan[i][0], new_stats)
        caches.append(c)
        skips.append(o)
        o = o[:, :, ::2]
    o, c = _block_fwd(o, state, plan[n][0], new_stats)         # :88
    caches.append(c)
    ups = []
    for j in range(n):                                          # :91-96
        U = upsample_matrix(o.shape[2])
        up = np.einsum("bcm,lm->bcl", o, U, optimize=True)
        ups.append((U, o.shape[1]))
        o = np.concatenate([up, skips[n - 1 - j]], axis=1)
        o, c = _block_fwd(o, state, plan[n + 1 + j][0], new_stats)
        caches.append(c)
    cat = np.concatenate([o, x], axis=1)                        # :98
    wo = state["out.0.weight"].astype(np.float64)
    bo = state["out.0.bias"].astype(np.float64)
    y = np.tanh(np.einsum("bcl,oc->bol", cat, wo[:, :, 0]) + bo[None, :, None])   # :99
    return y, new_stats, (caches, ups, cat, wo, y, n, plan)


def backward(cache, dy: np.ndarray) -> Dict[str, np.ndarray]:
    """Parameter gradients for the upstream gradient dy [B,1,T] of the output."""
    caches, ups, cat, wo, y, n, plan = cache
    grads: Dict[str, np.ndarray] = {}
    dpre = dy * (1.0 - y * y)
    grads["out.0.weight"] = np.einsum("bol,bcl->oc", dpre, cat)[:, :, None]
    grads["out.0.bias"] = dpre.sum(axis=(0, 2))
    dcat = np.einsum("bol,oc->bcl", dpre, wo[:, :, 0])
    do = dcat[:, :-1]                                           # the raw-input channel needs no gradient
    dskips = [None] * n
    for j in reversed(range(n)):
        dcat_j = _block_bwd(do, caches[n + 1 + j], plan[n + 1 + j][0], grads)
        U, cprev = ups[j]
        dskips[n - 1 - j] = dcat_j[:, cprev:]
        do = np.einsum("bcl,lm->bcm", dcat_j[:, :cprev], U, optimize=True)   # adjoint of the interpolation
    do = _block_bwd(do, caches[n], plan[n][0], grads)
    for i in reversed(range(n)):
        da = dskips[i].copy()
        da[:, :, ::2] += do                                     # adjoint of o[:, :, ::2]
        do = _block_bwd(da, caches[i], plan[i][0], grads)
    return grads


def mse_step(state, x: np.ndarray, clean: np.ndarray, n_layers: int = 12, channels_interval: int = 24):
    """One reference training step up to the gradients: -> (loss, grads, new running stats, y)."""
    y, new_stats, cache = forward_train(state, x, n_layers, channels_interval)
    diff = y - clean.astype(np.float64)
    loss = float((diff ** 2).mean())
    grads = backward(cache, 2.0 * diff / diff.size)
    return loss, grads, new_stats, y
