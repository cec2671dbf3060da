"""bf16 tensor-core path against its arithmetic model (oracle/wunet_bf16_model.py): level by level, the kernels must
reproduce the model's bf16 values except where the fp32 accumulation order flips a rounding (a one-ulp difference at a
small fraction of the elements). Much sharper than the tolerance against the fp32 oracle (test_parity_gpu.py).

A value may differ from the model's by one bf16 ulp where the fp32 accumulation order flips the final rounding; where the
pre-activation cancels to almost zero the flip is an fp32 rounding error of the SUM (relative to the terms, not the result),
so differences are also accepted up to ABS_FLOOR x the level's largest magnitude."""
import os

import numpy as np
import pytest
import torch

from oracle import wunet_bf16_model as wb
from oracle import wunet_oracle as wo
from wave_u_net_for_speech_enhancement_b200 import Model

pytestmark = [pytest.mark.gpu]

ABS_FLOOR = 2.0 ** -11       # of the level's max |value|


def lib_enc0_tc(ci, T):
    """Does the library run the first block on the tensor cores (3 taps over groups of 8 samples, bf16 samples and weights) for
    this shape? Mirrors build_plan: WUNET_TC_ENC0, frames of at least 1024 samples, 8 C <= 256 columns."""
    return os.environ.get("WUNET_TC_ENC0", "0") == "1" and T % 8 == 0 and T // 8 >= 128 and ci % 8 == 0 and 8 * ci <= 256


def ulp_bf16(v):
    """spacing of bf16 numbers at |v| (8 significant bits)"""
    return np.exp2(np.floor(np.log2(np.maximum(np.abs(v), 1e-30))) - 7)


# the third case has batch >= 64: blocks of at most 16 samples (enc5, middle, dec0 there) run the dense GEMM over frames, the
# decoder among them with the interpolation folded into its weights
@pytest.mark.parametrize("n,ci,B,T,seed", [(4, 8, 3, 256, 11), (12, 24, 2, 16384, 0), (6, 8, 64, 512, 5)])
def test_levels_match_the_arithmetic_model(n, ci, B, T, seed, monkeypatch):
    monkeypatch.setenv("WUNET_TC_STORE_LAST", "1")               # materialise the last decoder block too
    st = wo.make_state(n, ci, seed=seed)
    x = wo.make_input(B, T, seed=seed + 100)
    m = Model(n, ci, precision="bf16")
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()})
    m = m.to("cuda:0").eval()
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda()).cpu().numpy()
    got_levels = [m.read_level(i, B, T).cpu().numpy() for i in range(2 * n + 1)]
    # every block of the model is evaluated on the GPU's own outputs of the blocks before it: a rounding flip stays local
    want_y, want = wb.forward_bf16_model(st, x, n, ci, return_levels=True, forced=got_levels[:2 * n],
                                            enc0_tc=lib_enc0_tc(ci, T))
    report = []
    for i in range(2 * n + 1):
        got = got_levels[i]
        ref = want[i] if i < 2 * n else want[i].astype(np.float32)
        if i == 2 * n:                                           # the stored copy of the last block is rounded to bf16
            ref = torch.from_numpy(ref).to(torch.bfloat16).float().numpy()
        diff = np.abs(got - ref)
        frac_equal = float((diff == 0).mean())
        excess = float((diff - np.maximum(2.0 * ulp_bf16(ref), ABS_FLOOR * np.abs(ref).max())).max())
        report.append((i, round(frac_equal, 5), float(diff.max()), excess))
    yerr = float(np.abs(y - want_y).max())
    print("bf16 model (teacher-forced) vs kernels: (block, fraction bit-equal, max abs diff, excess over bound)")
    for r in report:
        print("   ", r)
    print("    output max-abs diff", yerr)
    assert all(r[1] >= 0.995 for r in report), report
    assert all(r[3] <= 0.0 for r in report), report
    assert yerr <= 2e-5, (yerr, report)
