"""Try a list of WUNET_TC_OVR overrides (one block each) and print that block's time next to the default plan's.
    timeout 120 python tools/ovr_try.py "23:mt=2,na=4" "22:mt=2,na=4" ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from wave_u_net_for_speech_enhancement_b200 import Model  # noqa: E402

torch.manual_seed(0)
m = Model(12, 24, precision="bf16").cuda().eval()
x = 0.3 * torch.randn(256, 1, 16384, device="cuda")


def times():
    with torch.no_grad():
        y = m(x)
        m.profile(True)
        tot = None
        for _ in range(5):
            y = m(x)
            ms = np.array(m.profile_read())
            tot = ms if tot is None else tot + ms
        m.profile(False)
    return y.clone(), tot / 5 * 1e3


os.environ.pop("WUNET_TC_OVR", None)
y0, t0 = times()
print("default: " + " ".join("%d:%.0f" % (i, v) for i, v in enumerate(t0) if i < 25), flush=True)
for ovr in sys.argv[1:]:
    os.environ["WUNET_TC_OVR"] = ovr
    try:
        y, t = times()
        blk = int(ovr.split(":")[0])
        print("%-28s block %2d: %.0f us (default %.0f)  exact=%s" % (ovr, blk, t[blk], t0[blk], bool(torch.equal(y, y0))), flush=True)
    except Exception as e:  # noqa: BLE001
        print("%-28s failed: %s" % (ovr, str(e)[:120]), flush=True)
os.environ.pop("WUNET_TC_OVR", None)
