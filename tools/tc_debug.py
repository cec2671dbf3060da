import os, sys
os.environ.setdefault("WUNET_TC_STORE_LAST", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import wunet_oracle as wo
from wave_u_net_for_speech_enhancement_b200 import Model
n, ci, T, B = 12, 24, 16384, 2
st = wo.make_state(n, ci, seed=0); x = wo.make_input(B, T, seed=1234)
y, levels = wo.COracle(n, ci).forward(st, x, return_levels=True)
m = Model(n, ci, precision="bf16")
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()})
m = m.to("cuda:0").eval()
with torch.no_grad():
    m(torch.from_numpy(x).cuda())
torch.cuda.synchronize()
for lvl in (10, 11):
    lv = m.read_level(lvl, B, T).cpu().numpy(); ref = levels[lvl]
    e = np.abs(lv - ref).max(axis=(0, 2))
    print("level", lvl, "per-channel max err (x1000):")
    print(np.round(e * 1000).astype(int).tolist())
    eb = np.abs(lv - ref).max(axis=(1,))
    print(" per (b,l) max err x1000:", np.round(eb * 1000).astype(int).tolist())
