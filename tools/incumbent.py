"""Incumbent Blackwell number: PyTorch-eager of the reference's forward (cuDNN / ATen kernels,
model/unet_basic.py:77-100 restated op for op by Model._forward_torch_reference_semantics) on the same GPU, eval mode.

    python tools/incumbent.py [B ...]     -> one JSON line per (B, mode)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from wave_u_net_for_speech_enhancement_b200 import Model  # noqa: E402


def time_mode(m, x, mode, iters=10, warm=3):
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"

    def run():
        with torch.no_grad():
            if mode == "bf16_autocast":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return m._forward_torch_reference_semantics(x)
            return m._forward_torch_reference_semantics(x)

    for _ in range(warm):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    Bs = [int(a) for a in sys.argv[1:]] or [1, 64, 256]
    torch.manual_seed(0)
    m = Model(12, 24).cuda().eval()
    for B in Bs:
        x = 0.3 * torch.randn(B, 1, 16384, device="cuda")
        for mode in ("fp32", "tf32", "bf16_autocast"):
            best, med = time_mode(m, x, mode)
            print(json.dumps({"incumbent": "torch-eager (cuDNN)", "B": B, "mode": mode, "best_ms": round(best, 4), "median_ms": round(med, 4),
                              "frames_per_s_best": round(B / best * 1e3, 1), "frames_per_s_median": round(B / med * 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
