"""bf16 path against the C oracle for channel plans other than the reference's 24 (block 1 runs in row-pair form wherever its
frames are long enough):   timeout 60 python tools/ci_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import wunet_oracle as wo  # noqa: E402
from wave_u_net_for_speech_enhancement_b200 import Model  # noqa: E402

for n, ci, B, T in [(5, 32, 3, 4096), (5, 16, 2, 4096), (6, 8, 3, 2048), (5, 32, 130, 1024)]:
    st = wo.make_state(n, ci, seed=1)
    x = wo.make_input(B, T, seed=2)
    want, lv = wo.COracle(n, ci).forward(st, x, return_levels=True)
    os.environ["WUNET_TC_STORE_LAST"] = "1"
    m = Model(n, ci, precision="bf16")
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()})
    m = m.cuda().eval()
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda()).cpu().numpy()
    b1 = m.read_level(1, B, T).cpu().numpy()
    print(f"n={n} ci={ci} B={B} T={T}: out err {np.abs(y - want).max():.3e}  block 1 rel err {np.abs(b1 - lv[1]).max() / np.abs(lv[1]).max():.4f}", flush=True)
    m._release()
