"""Per-block times of the bf16 forward (B=256) with the library given by WUNET_LIB_PATH (A/B of development builds):

    for so in build/libw_*.so; do WUNET_LIB_PATH=$so timeout 120 python tools/lib_times.py; done
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from wave_u_net_for_speech_enhancement_b200 import Model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
torch.manual_seed(0)
m = Model(12, 24, precision=prec).cuda().eval()
x = 0.3 * torch.randn(B, 1, 16384, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
with torch.no_grad():
    for _ in range(3):
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = m(x)
    e1.record()
    torch.cuda.synchronize()
    whole = e0.elapsed_time(e1) / 20
    m.profile(True)
    tot = None
    for _ in range(5):
        y = m(x)
        ms = np.array(m.profile_read())
        tot = ms if tot is None else tot + ms
tot = tot / 5
print("%s B=%d %s: %.4f ms/forward (events around 20), sum of blocks %.4f ms, checksum %.6f" %
      (os.path.basename(os.environ.get("WUNET_LIB_PATH", "libwunet_b200.so")), B, prec, whole, tot.sum(), float(y.double().abs().sum())))
print("   per block us: " + " ".join("%d:%.0f" % (i, a * 1e3) for i, a in enumerate(tot) if i < 25), flush=True)
m._release()
