"""Training-step oracle (row N1 of SURVEY.md §8f, groundwork): the numpy float64 restatement of forward(train-mode
BatchNorm) + backward in oracle/wunet_train_oracle.py against the vectors produced by autograd on the live reference
module (oracle/gen_golden_train.py). CPU only."""
import os

import numpy as np
import pytest

from oracle import wunet_oracle as wo
from oracle import wunet_train_oracle as wt

REL = 1e-9          # both sides are float64; differences are summation order only


def make_pair(B, T, seed):
    g = np.random.Generator(np.random.PCG64(seed))
    clean = (0.1 * g.standard_normal((B, 1, T))).astype(np.float32)
    noisy = (clean + 0.05 * g.standard_normal((B, 1, T))).astype(np.float32)
    return noisy, clean


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_small_config_every_gradient(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_small_n4_c8.npz"))
    n, ci, B, T = int(g["n_layers"]), int(g["channels_interval"]), int(g["B"]), int(g["T"])
    st = wo.make_state(n, ci, seed=int(g["state_seed"]))
    noisy, clean = make_pair(B, T, int(g["pair_seed"]))
    loss, grads, stats, y = wt.mse_step(st, noisy, clean, n, ci)
    assert abs(loss - float(g["loss"])) <= REL * abs(float(g["loss"]))
    assert rel_err(y, g["y"]) <= REL
    keys = [k[5:] for k in g.files if k.startswith("grad:")]
    assert sorted(keys) == sorted(grads.keys()) and len(keys) == 4 * (2 * n + 1) + 2
    for k in keys:
        assert grads[k].shape == g["grad:" + k].shape, k
        if k.endswith(".0.bias") and not k.startswith("out."):
            # a conv bias in front of a training-mode BatchNorm has an exactly zero gradient: rounding noise on both sides
            assert np.abs(grads[k]).max() <= 1e-12 and np.abs(g["grad:" + k]).max() <= 1e-12, k
            continue
        assert rel_err(grads[k], g["grad:" + k]) <= 1e-7, k
    for k in [k[5:] for k in g.files if k.startswith("stat:")]:
        if "num_batches" in k:
            continue
        assert rel_err(stats[k], g["stat:" + k]) <= REL, k


def test_reference_architecture_short_frames(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_full_n12_c24_b2_t4096.npz"))
    n, ci, B, T = int(g["n_layers"]), int(g["channels_interval"]), int(g["B"]), int(g["T"])
    st = wo.make_state(n, ci, seed=int(g["state_seed"]))
    noisy, clean = make_pair(B, T, int(g["pair_seed"]))
    loss, grads, stats, y = wt.mse_step(st, noisy, clean, n, ci)
    assert abs(loss - float(g["loss"])) <= REL * abs(float(g["loss"]))
    assert rel_err(y, g["y"]) <= REL
    for k, v in grads.items():
        if k.endswith(".0.bias") and not k.startswith("out."):
            continue                                            # exactly cancelled by BatchNorm: pure rounding noise on both sides
        norm = float(g["gnorm:" + k])
        assert abs(np.sqrt((v ** 2).sum()) - norm) <= 1e-7 * norm, k
        assert np.abs(v.reshape(-1)[g["gidx:" + k]] - g["gval:" + k]).max() <= 1e-7 * max(np.abs(v).max(), 1e-30), k
    for k in [k[5:] for k in g.files if k.startswith("stat:")]:
        if "num_batches" in k:
            continue
        assert rel_err(stats[k], g["stat:" + k]) <= REL, k


def test_upsample_matrix_is_the_forward_oracles_interpolation():
    x = np.random.default_rng(0).standard_normal((2, 3, 16))
    U = wt.upsample_matrix(16)
    # the forward oracle does the index math in fp32 like ATen's float kernel, this matrix in float64 like the .double() reference
    assert np.abs(np.einsum("bcm,lm->bcl", x, U) - wo.upsample_linear_x2_np(x)).max() <= 1e-5
    assert np.allclose(U.sum(axis=1), 1.0)


def test_gradient_check_by_finite_differences():
    """Independent of the goldens: d loss / d theta by central differences on a tiny network."""
    n, ci, B, T = 2, 8, 2, 16
    st = {k: v.astype(np.float64) for k, v in wo.make_state(n, ci, seed=5).items()}
    noisy, clean = make_pair(B, T, 7)
    _, grads, _, _ = wt.mse_step(st, noisy, clean, n, ci)
    rng = np.random.default_rng(3)
    for key in ["encoder.0.main.0.weight", "encoder.1.main.1.weight", "middle.0.weight", "decoder.0.main.0.weight",
                "decoder.1.main.1.bias", "out.0.weight", "out.0.bias"]:
        flat_idx = rng.integers(0, st[key].size, size=3)
        for fi in flat_idx:
            idx = np.unravel_index(fi, st[key].shape)
            h = 1e-6
            sp = {k: v.copy() for k, v in st.items()}
            sm = {k: v.copy() for k, v in st.items()}
            sp[key][idx] += h
            sm[key][idx] -= h
            lp = wt.mse_step(sp, noisy, clean, n, ci)[0]
            lm = wt.mse_step(sm, noisy, clean, n, ci)[0]
            fd = (lp - lm) / (2 * h)
            assert abs(fd - grads[key][idx]) <= 1e-6 * max(1.0, abs(fd)) + 1e-9, (key, idx, fd, grads[key][idx])
