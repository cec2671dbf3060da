"""Data-parallel training step on real GPUs (BASELINE.json configs[4]): one process per GPU over NCCL.
The native backward averages the gradients inside loss.backward() (flat bucket, two overlapped parts); this checks it against
the same step with the in-backward reduction switched off followed by an explicit all-reduce, and that the ranks stay in
lock-step after Adam. Needs >= 2 visible GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_ddp_gpu.py -m gpu`)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["WUNET_ROOT"])
import numpy as np, torch, torch.distributed as dist
from oracle import wunet_oracle as wo
from wave_u_net_for_speech_enhancement_b200 import Model
from wave_u_net_for_speech_enhancement_b200.train_step import allreduce_gradients, train_step
rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
n, ci, B, T = 12, 24, 4, 4096
st = {k: torch.from_numpy(np.asarray(v)) for k, v in wo.make_state(n, ci, seed=2).items()}
g = np.random.Generator(np.random.PCG64(50 + rank))                     # every rank its own shard of the batch
clean = torch.from_numpy((0.1 * g.standard_normal((B, 1, T))).astype(np.float32)).to(dev)
noisy = clean + torch.from_numpy((0.05 * g.standard_normal((B, 1, T))).astype(np.float32)).to(dev)
def grads(data_parallel):
    m = Model(n, ci, data_parallel=data_parallel)
    m.load_state_dict(st)
    m = m.to(dev).train()
    loss = torch.nn.MSELoss()(clean, m(noisy))
    loss.backward()
    return m
m_auto = grads("auto")
assert m_auto.reduces_gradients
m_off = grads("off")
assert not m_off.reduces_gradients
allreduce_gradients(m_off.parameters())                                  # explicit reduction of the local gradients
worst = 0.0
for (k, a), b in zip(m_auto.named_parameters(), m_off.parameters()):
    d = float((a.grad - b.grad).abs().max()) / max(float(b.grad.abs().max()), 1e-30)
    worst = max(worst, d)
assert worst <= 1e-4, worst          # the weight-gradient slices are combined with atomicAdd: run-to-run order noise ~2e-6
flat = torch.cat([p.grad.reshape(-1) for p in m_auto.parameters()])
gathered = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert all(torch.equal(gathered[0], t) for t in gathered), "ranks hold different gradients after backward()"
# three optimizer steps through the public step: replicas must stay identical
m = Model(n, ci); m.load_state_dict(st); m = m.to(dev).train()
opt = torch.optim.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.999))
for _ in range(3):
    train_step(m, opt, torch.nn.MSELoss(), noisy, clean)
flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
gathered = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert all(torch.equal(gathered[0], t) for t in gathered), "replicas diverged"
if rank == 0:
    print("DDP_OK worst relative difference in-backward vs explicit all-reduce: %.2e" % worst)
dist.destroy_process_group()
'''


def test_two_rank_native_training_step(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, WUNET_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "DDP_OK" in out.stdout
    print(out.stdout.strip().splitlines()[-1])
