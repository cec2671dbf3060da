"""A/B of the merged decoder tail chunk (WUNET_TC_MERGE=0 switches it off): output difference and per-block times.

    timeout 120 python tools/merge_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from wave_u_net_for_speech_enhancement_b200 import Model  # noqa: E402


def run(merge, B, reps=5):
    os.environ["WUNET_TC_MERGE"] = "1" if merge else "0"      # read when the library creates the model's tensor-core state
    torch.manual_seed(0)
    m = Model(12, 24, precision="bf16").cuda().eval()
    x = 0.3 * torch.randn(B, 1, 16384, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
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


for B in (3, 256):
    y0, t0 = run(False, B)
    y1, t1 = run(True, B)
    print("B=%d: unmerged %.4f ms, merged %.4f ms, max|diff| %.3e" % (B, t0.sum(), t1.sum(), float((y1 - y0).abs().max())), flush=True)
    if B == 256:
        print("   per block us: " + " ".join("%d:%.0f>%.0f" % (i, a * 1e3, b * 1e3) for i, (a, b) in enumerate(zip(t0, t1)) if i < 25))
