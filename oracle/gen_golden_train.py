"""
oracle/gen_golden_train.py — generates tests/golden/train_*.npz by running ONE training step of the LIVE reference.

Run in the authoring container only (needs /root/reference):

    python oracle/gen_golden_train.py

The unmodified ``model.unet_basic.Model`` (model/unet_basic.py:32) is cast to float64, put in ``.train()`` mode, fed the
seeded synthetic pair of SURVEY.md §8d config 5 (clean = 0.1·randn, noisy = clean + 0.05·randn, numpy PCG64), and stepped
as ``trainer/trainer.py:34-37`` does: forward, ``nn.MSELoss()`` (model/loss.py:3-4), ``backward()``.  Stored: the loss,
the output, every parameter gradient (small config) or its norm and probes (full config), and the updated BatchNorm
running statistics.  These vectors pin ``oracle/wunet_train_oracle.py`` (tests/test_train_oracle.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("WUNET_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from model.unet_basic import Model  # noqa: E402  (the unmodified reference)
from model.loss import mse_loss  # noqa: E402
from oracle import wunet_oracle as wo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make_pair(B, T, seed):
    g = np.random.Generator(np.random.PCG64(seed))
    clean = (0.1 * g.standard_normal((B, 1, T))).astype(np.float32)
    noisy = (clean + 0.05 * g.standard_normal((B, 1, T))).astype(np.float32)
    return noisy, clean


def step(n, ci, state, noisy, clean):
    m = Model(n_layers=n, channels_interval=ci)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=True)
    m = m.double().train()
    y = m(torch.from_numpy(noisy).double())
    loss = mse_loss()(torch.from_numpy(clean).double(), y)
    loss.backward()
    grads = {k: p.grad.detach().numpy() for k, p in m.named_parameters()}
    stats = {k: v.detach().numpy() for k, v in m.state_dict().items() if "running_" in k}
    return float(loss.detach()), y.detach().numpy(), grads, stats


def probes(a, n=16):
    flat = a.reshape(-1)
    idx = np.unique(np.linspace(0, flat.size - 1, num=min(n, flat.size)).round().astype(np.int64))
    return idx, flat[idx]


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(OUT, exist_ok=True)
    # ---- small config: every gradient stored ----------------------------------------------------------------
    n, ci, B, T = 4, 8, 3, 256
    st = wo.make_state(n, ci, seed=21)
    noisy, clean = make_pair(B, T, seed=22)
    loss, y, grads, stats = step(n, ci, st, noisy, clean)
    np.savez_compressed(os.path.join(OUT, "train_small_n4_c8.npz"), n_layers=n, channels_interval=ci, B=B, T=T,
                        state_seed=21, pair_seed=22, loss=loss, y=y,
                        **{"grad:" + k: v for k, v in grads.items()}, **{"stat:" + k: v for k, v in stats.items()})
    # ---- reference architecture (12 x 24), short frames: norms + probes ----------------------------------------
    n, ci, B, T = 12, 24, 2, 4096
    st = wo.make_state(n, ci, seed=0)
    noisy, clean = make_pair(B, T, seed=23)
    loss, y, grads, stats = step(n, ci, st, noisy, clean)
    out = dict(n_layers=n, channels_interval=ci, B=B, T=T, state_seed=0, pair_seed=23, loss=loss, y=y)
    for k, v in grads.items():
        idx, val = probes(v)
        out["gnorm:" + k] = np.sqrt((v.astype(np.float64) ** 2).sum())
        out["gidx:" + k] = idx
        out["gval:" + k] = val
    for k, v in stats.items():
        out["stat:" + k] = v
    np.savez_compressed(os.path.join(OUT, "train_full_n12_c24_b2_t4096.npz"), **out)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.startswith("train_")))


if __name__ == "__main__":
    main()
