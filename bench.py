#!/usr/bin/env python
"""
bench.py — 16384-sample frames/s of the Wave-U-Net forward (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--precision fp32|bf16|auto]
    python bench.py --impl reference ...        # the reference's CPU implementation of the path (host cores)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W   # one rank per GPU, batch sharded, no collective on the path

A "step" is one forward over one batch of B synthetic frames per GPU ([B,1,16384] fp32, eval mode).
Workload at N=1 = BASELINE.json configs[2] shape (batch 256, 16384 samples, 12 levels); weak scaling:
every rank runs the same B.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_LAYERS, CH_INT, T = 12, 24, 16384
METRIC = "16384-sample frames/sec forward"


# ------------------------------------------------------------------------------------------------
# roofline model: algorithmic flops / bytes per block for a batch of B frames
# ------------------------------------------------------------------------------------------------
def block_table(B: int, esize: int):
    """[(name, flops, bytes)] for the 2n+1 conv blocks (+ the head folded into the last decoder).
    bytes = fused-minimum traffic: read inputs once, write the output once, weights once per batch."""
    n, ci = N_LAYERS, CH_INT
    rows = []
    enc_out = [(i + 1) * ci for i in range(n)]
    for i in range(n):
        cin = 1 if i == 0 else i * ci
        cout = enc_out[i]
        L = T >> i
        flops = 2.0 * cin * cout * 15 * L * B
        in_b = (4 if i == 0 else esize) * cin * L            # raw fp32 input for enc0, else even half of skip_{i-1}
        byts = B * (in_b + esize * cout * L) + esize * cin * cout * 15
        rows.append((f"enc{i}", flops, byts))
    L = T >> n
    c = n * ci
    rows.append(("middle", 2.0 * c * c * 15 * L * B, B * 2 * esize * c * L + esize * c * c * 15))
    cprev = c
    for j in range(n):
        e = n - 1 - j
        cskip = enc_out[e]
        cout = enc_out[e]
        L = T >> e
        cin = cprev + cskip
        flops = 2.0 * cin * cout * 5 * L * B
        reads = esize * (cprev * (L // 2) + cskip * L)
        writes = esize * cout * L
        if j == n - 1:                                        # head fused: + raw input read, 1-channel fp32 write
            flops += 2.0 * (cout + 1) * L * B
            reads += 4 * L
            writes = 4 * L
        rows.append((f"dec{j}" + ("+out" if j == n - 1 else ""), flops, B * (reads + writes) + esize * cin * cout * 5))
        cprev = cout
    return rows


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tensor_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json; bf16 sustained, kernels timed inside a long step)"}
    return {"hbm_gbs": 6650.0, "tensor_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference forward restated with torch ops (oracle/wunet_oracle.torch_port_forward)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, frames_per_step: int, budget_s: float):
    from oracle import wunet_oracle as wo
    st = {k: torch.from_numpy(np.asarray(v)) for k, v in wo.make_state(N_LAYERS, CH_INT, seed=0).items()}
    x = torch.from_numpy(wo.make_input(frames_per_step, T, seed=1234))
    threads = torch.get_num_threads()
    with torch.no_grad():
        for _ in range(max(1, warmup)):
            wo.torch_port_forward(st, x)
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            wo.torch_port_forward(st, x)
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return {"value": frames_per_step * done / dt, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{done} forwards of B={frames_per_step} x 16384 samples, fp32, torch CPU ops "
                      f"(oneDNN) restating model/unet_basic.py:77-100, {threads} threads of {os.cpu_count()} cpus",
            "ms_per_step": 1e3 * dt / done, "steps_done": done}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner, ...) write to fd 1; keep the real stdout for the ONE JSON line only
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    # --------------------------------------------------------------------------------------------
    # reference arm: the reference's CPU implementation of the path, on the host cores (rank 0 only)
    # --------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        fps = 8
        r = cpu_reference_run(steps=args.steps, warmup=args.warmup, frames_per_step=fps, budget_s=120.0)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "frames/s", "n_gpus": args.gpus,
                "steps": r["steps_done"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"unet_basic forward, {N_LAYERS} levels, {CH_INT} base filters, 16384 samples, "
                                       f"bounded sample of {fps} frames/step on CPU"},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # --------------------------------------------------------------------------------------------
    # B200 arm
    # --------------------------------------------------------------------------------------------
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback on the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from oracle import wunet_oracle as wo                    # synthetic weights/inputs + cpu_baseline only
    from wave_u_net_for_speech_enhancement_b200 import Model, _lib

    B = args.batch
    precision = args.precision
    if precision == "auto":
        precision = "bf16"
    st = {k: torch.from_numpy(np.asarray(v)) for k, v in wo.make_state(N_LAYERS, CH_INT, seed=0).items()}

    def make(prec):
        m = Model(N_LAYERS, CH_INT, precision=prec)
        m.load_state_dict(st, strict=True)
        return m.to(dev).eval()

    model = make(precision)
    # rotating device-resident inputs (the per-step working set, >3 GB of activations, is far larger than L2)
    NBUF = 4
    xs = [torch.from_numpy(wo.make_input(B, T, seed=1234 + 17 * rank + i)).to(dev) for i in range(NBUF)]
    try:
        with torch.no_grad():
            model(xs[0])
        torch.cuda.synchronize()
    except _lib.WunetError as e:
        if args.precision == "auto" and "not built yet" in str(e):
            precision = "fp32"
            model = make(precision)
        else:
            raise

    def step(i):
        with torch.no_grad():
            return model(xs[i % NBUF])

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    launches_per_step = model.last_launch_count()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.start()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    # the timed region is tens of milliseconds; nvidia-smi samples every 100 ms. Keep issuing the identical step (untimed)
    # for ~1.2 s so that the clock / throttle record is taken under the same load.
    t_load = time.perf_counter()
    extra = 0
    while time.perf_counter() - t_load < 1.2:
        for i in range(args.steps):
            step(i)
        torch.cuda.synchronize()
        extra += args.steps
    barrier()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["note"] = f"sampled every 100 ms over the timed steps plus {extra} identical untimed steps"
    t = torch.tensor([ms_total], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total / 1e3)

    # ---- per-level device times (CUDA events on the launch stream, same K steps) -> roofline ------------
    model.profile(True)
    per_block = None
    for i in range(args.steps):
        step(i)
        ms = np.array(model.profile_read())
        per_block = ms if per_block is None else per_block + ms
    model.profile(False)
    per_block = per_block / args.steps
    peaks = load_peaks()
    esize = 2 if precision == "bf16" else 4
    tab = block_table(B, esize)
    blk_ms = list(per_block[:len(tab)])
    if len(per_block) > len(tab):
        blk_ms[-1] += float(per_block[len(tab)])              # separate head launch belongs to dec11+out
    levels = []
    for (name, flops, byts), t_ms in zip(tab, blk_ms):
        t_mem = byts / (peaks["hbm_gbs"] * 1e9) * 1e3
        t_ten = flops / (peaks["tensor_tflops"] * 1e12) * 1e3
        bound = "hbm" if t_mem >= t_ten else "tensor"
        levels.append({"level": name, "ms": round(float(t_ms), 5), "roof_ms": round(max(t_mem, t_ten), 5), "bound": bound,
                       "gbs": round(byts / (t_ms * 1e-3) / 1e9, 1), "tflops": round(flops / (t_ms * 1e-3) / 1e12, 2),
                       "frac": round(max(t_mem, t_ten) / float(t_ms), 4)})
    dom = max(range(len(tab)), key=lambda i: blk_ms[i])
    dname, dflops, dbytes = tab[dom]
    d = levels[dom]
    if d["bound"] == "hbm":
        roof = {"bound": "hbm", "achieved": d["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": round(d["gbs"] / peaks["hbm_gbs"], 4)}
    else:
        roof = {"bound": "tensor", "achieved": d["tflops"], "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
                "frac": round(d["tflops"] / peaks["tensor_tflops"], 4)}
    traffic = None
    try:
        import csv, glob
        cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_conv_tc_*summary.csv")))
        if cands and precision == "bf16" and B == 256:
            rows = [r for r in csv.reader(open(cands[-1])) if r and not r[0].startswith("#")]
            hdr = rows[0]
            for r in rows[1:]:
                if r[0] == dname.split("+")[0]:
                    traffic = (float(r[hdr.index("dram_rd_MB")]) + float(r[hdr.index("dram_wr_MB")])) * 1e6
    except Exception:
        traffic = None
    roof.update({"kernel": dname, "kernel_ms": d["ms"], "traffic": traffic, "algorithmic_bytes": dbytes,
                 "algorithmic_flops": dflops, "peak_source": peaks["source"],
                 "net_frac": round(sum(l["roof_ms"] for l in levels) / sum(l["ms"] for l in levels), 4),
                 "sum_levels_ms": round(float(sum(blk_ms)), 4)})

    # ---- end to end: public API with HOST buffers (pinned); every step copies ITS inputs H2D and ITS result D2H inside
    # the timed region. (a) streaming API Model.forward_host_stream: two batches in flight, so the copies of step k+1 / k-1
    # overlap the kernels of step k (the way enhancement.py's many chunks are served); (b) one synchronous forward_host per step.
    NH = 4
    xh = [torch.from_numpy(wo.make_input(B, T, seed=99 + i)).pin_memory() for i in range(NH)]
    yh = [torch.empty_like(xh[0]).pin_memory() for _ in range(NH)]
    for i in range(2):
        model.forward_host(xh[i % NH], out=yh[i % NH])
    for _ in model.forward_host_stream([xh[i % NH] for i in range(3)], [yh[i % NH] for i in range(3)]):
        pass
    barrier()
    t0 = time.perf_counter()
    for _ in model.forward_host_stream((xh[i % NH] for i in range(args.steps)), (yh[i % NH] for i in range(args.steps))):
        pass
    torch.cuda.synchronize()
    dt_stream = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        model.forward_host(xh[i % NH], out=yh[i % NH])
    torch.cuda.synchronize()
    dt_sync = time.perf_counter() - t0
    t = torch.tensor([dt_stream, dt_sync], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = B * world * args.steps / float(t[0].item())
    e2e_sync = B * world * args.steps / float(t[1].item())
    nbytes = B * T * 4

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(steps=1000, warmup=1, frames_per_step=8, budget_s=15.0)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"unet_basic forward, batch {B}/GPU x 16384 samples, {N_LAYERS} levels, {CH_INT} base "
                               f"filters, eval-mode BatchNorm, {precision} path",
                   "global_batch": B * world, "parallelism": f"dp{world} (batch sharded, no collective on the forward)",
                   "l2": f"{NBUF} rotating input batches; per-step activation working set "
                         f"{sum(b for _n, _f, b in tab) / 2**30:.2f} GiB >> 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                "api": "Model.forward_host_stream -> wunet_stream_submit/wait: pinned host buffers in and out every step, "
                       "two batches in flight",
                "sync_value": e2e_sync, "sync_api": "Model.forward_host -> wunet_forward_host (one blocking call per step)"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roof,
        "levels": levels,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
