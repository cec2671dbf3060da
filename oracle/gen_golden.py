"""
oracle/gen_golden.py — generates tests/golden/* by running the LIVE reference module.

Run in the authoring container only (needs /root/reference):

    python oracle/gen_golden.py

It imports ``model.unet_basic.Model`` unchanged from ``/root/reference`` (model/unet_basic.py:32),
loads the seeded synthetic weights of ``oracle.wunet_oracle.make_state`` with strict
``load_state_dict``, runs ``Model.forward`` (model/unet_basic.py:77-100) in eval mode on CPU fp32
(``CUDA_VISIBLE_DEVICES=-1`` semantics of enhancement.py:25) and stores outputs plus per-level
intermediates (forward hooks on encoder[i] / middle / decoder[j]).  The vectors pin both the oracle
(tests/test_oracle.py) and the CUDA path (tests/test_parity_gpu.py); /root/reference itself does not
travel to the GPU box.
"""
import json
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
from oracle import wunet_oracle as wo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run_reference(n, ci, state, x):
    m = Model(n_layers=n, channels_interval=ci)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}
    m.load_state_dict(sd, strict=True)
    m.eval()
    levels = []
    hooks = []
    for mod in list(m.encoder) + [m.middle] + list(m.decoder):
        hooks.append(mod.register_forward_hook(lambda _m, _i, o: levels.append(o.detach().clone().numpy())))
    with torch.no_grad():
        y = m(torch.from_numpy(x)).numpy()
    for h in hooks:
        h.remove()
    return y, levels, m


def probe_index(L, n=24):
    idx = set(range(min(L, 8))) | set(range(max(0, L - 8), L))
    idx |= set(int(v) for v in np.linspace(0, L - 1, num=min(L, n - 16) if L > 16 else L).round())
    return np.array(sorted(idx), dtype=np.int64)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(OUT, exist_ok=True)

    # ---- state_dict surface of the reference -------------------------------------
    m = Model()
    surface = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()]
    params = [[k, list(v.shape)] for k, v in m.named_parameters()]
    with open(os.path.join(OUT, "state_dict_surface.json"), "w") as f:
        json.dump({"state_dict": surface, "named_parameters": params,
                   "n_params": int(sum(p.numel() for p in m.parameters()))}, f, indent=0)
    # the oracle's own key list must be the reference's
    assert [(k, tuple(s), d) for k, s, d in surface] == wo.state_keys(12, 24), "state_keys mismatch"

    # ---- small config: everything stored --------------------------------------------------------
    n, ci, T, B = 4, 8, 256, 3
    st = wo.make_state(n, ci, seed=11)
    x = wo.make_input(B, T, seed=12)
    y, levels, _ = run_reference(n, ci, st, x)
    np.savez_compressed(os.path.join(OUT, "small_n4_c8.npz"), n_layers=n, channels_interval=ci, T=T, B=B,
                        state_seed=11, input_seed=12, y=y,
                        **{f"level_{i}": l for i, l in enumerate(levels)})

    # ---- full config (12 levels, 24 base filters, T=16384), B=2 -----------------------------------
    n, ci, T, B = 12, 24, 16384, 2
    st = wo.make_state(n, ci, seed=0)
    x = wo.make_input(B, T, seed=1234)
    y, levels, _ = run_reference(n, ci, st, x)
    full = dict(n_layers=n, channels_interval=ci, T=T, B=B, state_seed=0, input_seed=1234, y=y)
    for i, l in enumerate(levels):
        idx = probe_index(l.shape[-1])
        full[f"probe_idx_{i}"] = idx
        full[f"probe_{i}"] = l[0][:, idx]                      # batch item 0, all channels
        full[f"sum_{i}"] = np.float64(l.astype(np.float64).sum())
        full[f"abssum_{i}"] = np.float64(np.abs(l.astype(np.float64)).sum())
    np.savez_compressed(os.path.join(OUT, "full_n12_c24_b2.npz"), **full)

    # ---- edge vectors on the full config ---------------------------------------------
    edges = {}
    for name, xe in wo.edge_inputs(T).items():
        ye, _, _ = run_reference(n, ci, st, xe)
        edges[name] = ye
    for Tx in (4096, 20480):                                     # non-default lengths work (T % 4096 == 0)
        xe = wo.make_input(1, Tx, seed=77 + Tx)
        ye, _, _ = run_reference(n, ci, st, xe)
        edges[f"T{Tx}"] = ye
    np.savez_compressed(os.path.join(OUT, "edges_n12_c24.npz"), state_seed=0, **edges)

    # T not a multiple of 2**n raises in the reference (torch.cat size error, unet_basic.py:95)
    try:
        run_reference(n, ci, st, wo.make_input(1, 16000, seed=5))
        raised = False
    except RuntimeError:
        raised = True
    assert raised, "reference accepted T=16000?"
    print("golden vectors written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print(f"  {fn}: {os.path.getsize(os.path.join(OUT, fn)) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
