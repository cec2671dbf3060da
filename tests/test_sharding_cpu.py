"""Host-side logic of the multi-GPU path on CPU (gloo, world_size 2): contiguous batch sharding, max-over-ranks timing
reduction and rank-0 reporting as bench.py does them. The forward itself needs no collective."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["WUNET_ROOT"])
import numpy as np, torch, torch.distributed as dist
from oracle import wunet_oracle as wo
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, ci, T, B = 3, 8, 64, 6                                   # global batch 6 -> 3 frames per rank
st = wo.make_state(n, ci, seed=1)
x = wo.make_input(B, T, seed=2)
shard = x[rank * (B // world):(rank + 1) * (B // world)]    # contiguous equal shards
y_local = wo.COracle(n, ci).forward(st, shard)              # stand-in for the per-rank forward (frames are independent)
t = torch.tensor([float(10 + rank)])                        # per-rank "elapsed ms"
dist.all_reduce(t, op=dist.ReduceOp.MAX)
gathered = [torch.zeros_like(torch.from_numpy(y_local)) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(y_local))        # test-only gather to compare with the unsharded result
# config 4 (enhancement.py:49-74 over several GPUs): clips dealt to the ranks by frame count, results gathered on rank 0
from wave_u_net_for_speech_enhancement_b200 import enhance
rng = np.random.default_rng(5)
lengths = [130, 64, 1, 700, 65, 333, 64, 250, 90]
clips = [rng.standard_normal(k).astype(np.float32) for k in lengths]
def fake_stream(batches, outs):
    for b_, o_ in zip(batches, outs):
        o_.copy_(b_ * 2.0 + torch.arange(b_.shape[-1], dtype=torch.float32) / b_.shape[-1])
        yield o_
shards = enhance.shard_clips(lengths, world, sample_length=64)
assert sorted(i for s_ in shards for i in s_) == list(range(len(lengths)))
fr = [sum(max(1, -(-lengths[i] // 64)) for i in s_) for s_ in shards]
assert max(fr) - min(fr) <= max(max(1, -(-k // 64)) for k in lengths), fr          # balanced to within one clip
res = enhance.enhance_waveforms_sharded(None, clips, rank, world, sample_length=64, batch_frames=4, stream_fn=fake_stream, gather=True)
if rank == 0:
    single = enhance.enhance_waveforms(None, clips, sample_length=64, batch_frames=4, stream_fn=fake_stream)
    assert len(res) == len(single) and all(np.array_equal(a, b_) for a, b_ in zip(res, single))
else:
    assert res is None
if rank == 0:
    y_full = wo.COracle(n, ci).forward(st, x)
    y_cat = torch.cat(gathered, 0).numpy()
    assert np.array_equal(y_cat, y_full), "sharded forward differs from the single-process forward"
    assert float(t) == 10 + world - 1
    print("SHARDING_OK")
dist.destroy_process_group()
'''


def test_two_rank_batch_sharding_matches_single_process(tmp_path):
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
    assert "SHARDING_OK" in out.stdout
