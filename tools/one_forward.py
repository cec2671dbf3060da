"""Run N forwards of the bf16 path at a given batch (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import wunet_oracle as wo
from wave_u_net_for_speech_enhancement_b200 import Model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
st = wo.make_state(12, 24, seed=0)
m = Model(12, 24, precision=prec)
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()})
m = m.cuda().eval()
x = torch.from_numpy(wo.make_input(B, 16384, seed=1)).cuda()
with torch.no_grad():
    for _ in range(reps):
        y = m(x)
torch.cuda.synchronize()
print("done", float(y.abs().max()))
m._release()
