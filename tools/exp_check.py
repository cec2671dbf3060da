"""Check and time the experimental kernel paths (WUNET_TC_EXP bit mask, X = 1 instantiations of conv_tc_kernel) against the
validated path. Run it under `timeout`: an experimental path that dead-locks must not take the GPU box with it.

    timeout 120 python tools/exp_check.py 1 2 3        # masks to try; prints max |diff| vs mask 0 and per-block times
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import wunet_oracle as wo  # noqa: E402
from wave_u_net_for_speech_enhancement_b200 import Model  # noqa: E402

st = wo.make_state(12, 24, seed=0)
sd = {k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}


def run(mask, B, reps=5):
    os.environ["WUNET_TC_EXP"] = str(mask)          # read when the library creates the model's tensor-core state
    m = Model(12, 24, precision="bf16")
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = torch.from_numpy(wo.make_input(B, 16384, seed=1)).cuda()
    with torch.no_grad():
        y = m(x)
        m.profile(True)
        tot = None
        for _ in range(reps):
            y = m(x)
            ms = np.array(m.profile_read())
            tot = ms if tot is None else tot + ms
    torch.cuda.synchronize()
    y = y.clone()
    m._release()
    return y, tot / reps


masks = [int(a) for a in sys.argv[1:]] or [1]
for B in (3, 256):
    y0, t0 = run(0, B)
    print("B=%d mask 0: %.4f ms" % (B, t0.sum()), flush=True)
    for mask in masks:
        y, t = run(mask, B)
        print("B=%d mask %d: %.4f ms, max|diff| %.3e, finite %s" % (B, mask, t.sum(), float((y - y0).abs().max()), bool(torch.isfinite(y).all())), flush=True)
        if B == 256:
            print("   per block us (mask 0 -> mask %d): " % mask + " ".join("%d:%.0f>%.0f" % (i, a * 1e3, b * 1e3) for i, (a, b) in enumerate(zip(t0, t)) if i < 25))
