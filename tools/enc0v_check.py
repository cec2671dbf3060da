"""A/B of the two CUDA-core forms of block 0 (WUNET_TC_ENC0V=1: one tile per block, 2: persistent): outputs must be bit-identical;
block 0's time at batch 256.   timeout 200 python tools/enc0v_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import wunet_oracle as wo  # noqa: E402
from wave_u_net_for_speech_enhancement_b200 import Model  # noqa: E402


def run(v, n, ci, B, T, prec, reps=0):
    os.environ["WUNET_TC_ENC0V"] = str(v)
    st = wo.make_state(n, ci, seed=3)
    m = Model(n, ci, precision=prec)
    m.load_state_dict({k: torch.from_numpy(np.asarray(a)) for k, a in st.items()})
    m = m.cuda().eval()
    x = torch.from_numpy(wo.make_input(B, T, seed=5)).cuda()
    with torch.no_grad():
        y = m(x).clone()
        b0 = m.read_level(0, B, T).clone()
        t0 = None
        if reps:
            m.profile(True)
            tot = 0.0
            for _ in range(reps):
                m(x)
                tot += m.profile_read()[0]
            m.profile(False)
            t0 = tot / reps * 1e3
            yh = m.forward_host(x.cpu().pin_memory()).clone()          # chunked host pipeline: enc0 launched per batch chunk
            assert torch.equal(yh, y.cpu()), "host pipeline differs"
    torch.cuda.synchronize()
    m._release()
    return y, b0, t0


for (n, ci, B, T, prec, reps) in [(12, 24, 3, 16384, "bf16", 0), (4, 8, 5, 2064, "bf16", 0), (6, 16, 2, 4160, "fp32_tc", 0),
                                   (12, 24, 64, 16384, "fp32_tc", 5), (12, 24, 256, 16384, "bf16", 10)]:
    y1, b1, t1 = run(1, n, ci, B, T, prec, reps)
    y2, b2, t2 = run(2, n, ci, B, T, prec, reps)
    print(f"n={n} ci={ci} B={B} T={T} {prec}: block0 equal {bool(torch.equal(b1, b2))} output equal {bool(torch.equal(y1, y2))}"
          + (f"  block 0: v1 {t1:.1f} us, v2 {t2:.1f} us" if t1 else ""), flush=True)
