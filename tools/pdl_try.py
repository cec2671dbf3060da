"""Step time of the bf16 forward (B=256) under combinations of WUNET_TC_PDL and WUNET_TC_OVR (both read when the library creates
the model's state / plan): does programmatic dependent launch pay once neighbouring blocks fit two CTAs per SM?
    timeout 200 python tools/pdl_try.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from wave_u_net_for_speech_enhancement_b200 import Model  # noqa: E402

x = 0.3 * torch.randn(256, 1, 16384, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
SMALL = "7:small=1;8:small=1;15:small=1"
y0 = None
for pdl, ovr in [("0", ""), ("1", ""), ("0", SMALL), ("1", SMALL), ("1", SMALL + ";18:small=1;20:small=1"), ("0", "")]:
    os.environ["WUNET_TC_PDL"] = pdl
    if ovr:
        os.environ["WUNET_TC_OVR"] = ovr
    else:
        os.environ.pop("WUNET_TC_OVR", None)
    torch.manual_seed(0)
    m = Model(12, 24, precision="bf16").cuda().eval()
    with torch.no_grad():
        for _ in range(3):
            y = m(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            y = m(x)
        e1.record()
        torch.cuda.synchronize()
        step = e0.elapsed_time(e1) / 30
        m.profile(True)
        tot = None
        for _ in range(5):
            m(x)
            ms = np.array(m.profile_read())
            tot = ms if tot is None else tot + ms
        m.profile(False)
    if y0 is None:
        y0 = y.clone()
    print("PDL=%s OVR=%-40s step %.4f ms  blocks 7..17 %.1f us  same output %s   [%s]" % (
        pdl, ovr, step, float(tot[7:18].sum() / 5 * 1e3), bool(torch.equal(y, y0)), " ".join("%.0f" % (v / 5 * 1e3) for v in tot[7:18])), flush=True)
    m._release()
