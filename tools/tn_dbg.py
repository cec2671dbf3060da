"""Development (library built with WUNET_TN_DEBUG=1): which role bounds the taps-in-N kernel? Times blocks 1 (enc1), 23 (dec10),
24 (dec11) with roles of the kernel switched off (WUNET_TN_DBG selects the instantiation at launch; results are numerically
meaningless). timeout 120 python tools/tn_dbg.py 0 1 2 4 8 16 32 3 7 24 56 63"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from wave_u_net_for_speech_enhancement_b200 import Model  # noqa: E402

t0 = time.time()
os.environ["WUNET_TC_TN"] = "1"
x = 0.3 * torch.randn(256, 1, 16384, device="cuda")
torch.manual_seed(0)
m = Model(12, 24, precision="bf16").cuda().eval()
with torch.no_grad():
    m(x)
    m.profile(True)
    for mask in [int(a) for a in sys.argv[1:]] or [0]:
        os.environ["WUNET_TN_DBG"] = str(mask)
        m(x)
        tot = None
        for _ in range(5):
            m(x)
            ms = np.array(m.profile_read())
            tot = ms if tot is None else tot + ms
        t = tot / 5 * 1e3
        print("dbg %2d: enc1 %.0f us, dec10 %.0f us, dec11 %.0f us | enc2 %.0f dec9 %.0f total %.0f  (%.1f s)" % (mask, t[1], t[23], t[24], t[2], t[22], t.sum(), time.time() - t0), flush=True)
