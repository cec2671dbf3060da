"""fp32_tc (split-precision tensor-core path) timing next to the bf16 and fp32 paths, and its error vs the fp32 path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from wave_u_net_for_speech_enhancement_b200 import Model  # noqa: E402

torch.manual_seed(0)
ref = Model(12, 24, precision="fp32").cuda().eval()
with torch.no_grad():
    g = torch.Generator().manual_seed(7)
    for mod in ref.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=g))
            mod.running_var.copy_(0.5 + torch.rand(mod.running_var.shape, generator=g))
            mod.weight.copy_(0.5 + torch.rand(mod.weight.shape, generator=g))
            mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
sd = ref.state_dict()
for B in (4, 64, 256):
    x = 0.3 * torch.randn(B, 1, 16384, device="cuda")
    with torch.no_grad():
        y32 = ref(x)
    for prec in ("fp32_tc", "bf16"):
        m = Model(12, 24, precision=prec)
        m.load_state_dict(sd)
        m = m.cuda().eval()
        with torch.no_grad():
            y = m(x)
            m.profile(True)
            tot = None
            for _ in range(5):
                m(x)
                ms = np.array(m.profile_read())
                tot = ms if tot is None else tot + ms
        t = tot / 5
        print("B=%d %s: %.3f ms/forward = %.0f frames/s, max|y - y_fp32| %.2e" % (B, prec, t.sum(), B / t.sum() * 1e3, float((y - y32).abs().max())), flush=True)
        if B == 256:
            print("   per block us: " + " ".join("%d:%.0f" % (i, v * 1e3) for i, v in enumerate(t) if i < 25), flush=True)
        m._release()
