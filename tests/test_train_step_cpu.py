"""Gradient all-reduce of the data-parallel training step (wave_u_net_for_speech_enhancement_b200/train_step.py) on CPU:
gloo, world_size 2. The model math itself is the training oracle's (numpy float64): each rank computes the gradients of
its shard, the all-reduce must leave both ranks with the mean over ranks, and an Adam step from it must be identical."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["WUNET_ROOT"])
import numpy as np, torch, torch.distributed as dist
from oracle import wunet_oracle as wo, wunet_train_oracle as wt
from wave_u_net_for_speech_enhancement_b200.train_step import allreduce_gradients
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, ci, T, B = 2, 8, 32, 4                                   # global batch 4 -> 2 frames per rank
st = wo.make_state(n, ci, seed=3)
g = np.random.Generator(np.random.PCG64(9))
clean = (0.1 * g.standard_normal((B, 1, T))).astype(np.float32)
noisy = (clean + 0.05 * g.standard_normal((B, 1, T))).astype(np.float32)
sl = slice(rank * (B // world), (rank + 1) * (B // world))
_, grads, _, _ = wt.mse_step(st, noisy[sl], clean[sl], n, ci)          # this rank's shard (per-shard BatchNorm statistics)
keys = sorted(grads)
params = [torch.nn.Parameter(torch.from_numpy(np.asarray(st[k], np.float64).copy())) for k in keys]
for p, k in zip(params, keys):
    p.grad = torch.from_numpy(grads[k].reshape(p.shape).copy())
allreduce_gradients(params)
# expectation: mean of the two ranks' shard gradients, computed locally from both shards
both = [wt.mse_step(st, noisy[r * 2:(r + 1) * 2], clean[r * 2:(r + 1) * 2], n, ci)[1] for r in range(world)]
for p, k in zip(params, keys):
    want = (both[0][k] + both[1][k]) / 2
    assert np.allclose(p.grad.numpy(), want.reshape(p.shape), rtol=1e-12, atol=1e-15), k
opt = torch.optim.Adam(params, lr=1e-3)
opt.step()
flat = torch.cat([p.detach().reshape(-1) for p in params])
gathered = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert torch.equal(gathered[0], gathered[1]), "ranks diverged after the optimizer step"
# the same through the public step (trainer/trainer.py:34-38 + all-reduce) with the composite backend of the drop-in model
from wave_u_net_for_speech_enhancement_b200 import Model
from wave_u_net_for_speech_enhancement_b200.train_step import train_step, GradientBucket
m = Model(n, ci, train_backend="torch")
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()})
m.train()
assert not m.reduces_gradients                               # only the native backend reduces inside backward()
opt = torch.optim.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.999))
loss = train_step(m, opt, lambda c, e: torch.nn.functional.mse_loss(e, c), torch.from_numpy(noisy[sl]), torch.from_numpy(clean[sl]))
flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
gathered = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert torch.equal(gathered[0], gathered[1]), "model replicas diverged after train_step"
# bucket layout: backward-completion order, two parts, every view aligned
b = GradientBucket([p.shape for p in m.parameters()], 2 * n + 1, "cpu")
names = [k for k, _ in m.named_parameters()]
first = [names[i] for i, v in enumerate(b.views) if v.storage_offset() < b.split]
assert all(k.startswith(("out.", "decoder.")) for k in first) and len(first) == 2 + 4 * n
assert all(v.storage_offset() % 4 == 0 for v in b.views) and b.split % 128 == 0
if rank == 0:
    print("ALLREDUCE_OK")
dist.destroy_process_group()
'''


def test_two_rank_gradient_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, WUNET_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ALLREDUCE_OK" in out.stdout
