"""
Chunk batching for inference: the reference's ``enhancement.py:49-74`` (and the validation loop,
``trainer/trainer.py:58-79``) zero-pads every clip to a multiple of ``sample_length``, splits it into 16384-sample chunks
and pushes them through the model ONE chunk at a time with a device→host sync each. Chunks are independent in eval mode,
so this module stacks the chunks of many clips on the batch axis and streams the batches through
``Model.forward_host_stream`` (H2D copy / kernels / D2H copy of consecutive batches overlap), then re-assembles and trims
every clip exactly like the reference loop does (``enhancement.py:68-71``).
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Sequence, Tuple

import numpy as np
import torch


class _PinnedPool:
    """Grow-only pinned staging buffers, reused across calls: allocating page-locked memory costs milliseconds, a forward of
    a few hundred frames costs one. Single-threaded use (every result is copied out of the pool before the call returns)."""

    def __init__(self):
        self._buf = {}

    def get(self, name: str, frames: int, sample_length: int, pin: bool) -> torch.Tensor:
        key = (name, sample_length, pin)
        t = self._buf.get(key)
        if t is None or t.shape[0] < frames:
            t = torch.empty(max(frames, 1), 1, sample_length, dtype=torch.float32)
            if pin:
                t = t.pin_memory()
            self._buf[key] = t
        return t[:frames]


_POOL = _PinnedPool()


def frame_clips(waveforms: Sequence[np.ndarray], sample_length: int = 16384, pin: bool = True
                ) -> Tuple[torch.Tensor, List[Tuple[int, int, int]]]:
    """Zero-pad each 1-D waveform to a multiple of ``sample_length`` (enhancement.py:57-59) and split it into chunks
    (enhancement.py:62). Returns the stacked chunks ``[N,1,sample_length]`` (float32; a view of a pooled pinned staging
    buffer if requested, valid until the next call) and, per clip, ``(first_frame, n_frames, original_length)``."""
    index, total = [], 0
    for w in waveforms:
        n = int(np.asarray(w).shape[-1])
        if np.asarray(w).ndim != 1:
            raise ValueError("waveforms must be 1-D")
        nf = max(1, -(-n // sample_length))
        index.append((total, nf, n))
        total += nf
    frames = _POOL.get("in", total, sample_length, pin and torch.cuda.is_available())
    flat = frames.view(total, sample_length)
    for (f0, nf, n), w in zip(index, waveforms):
        dst = flat[f0:f0 + nf].reshape(-1)
        dst[:n] = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
        dst[n:] = 0.0                                          # the padding of enhancement.py:58 (the buffer is reused)
    return frames, index


def unframe_clips(frames: torch.Tensor, index: List[Tuple[int, int, int]]) -> List[np.ndarray]:
    """Concatenate each clip's chunks and drop the padding (enhancement.py:68-71)."""
    flat = frames.view(frames.shape[0], frames.shape[-1])
    return [flat[f0:f0 + nf].reshape(-1)[:n].numpy().copy() for (f0, nf, n) in index]


def enhance_waveforms(model, waveforms: Sequence[np.ndarray], sample_length: int = 16384, batch_frames: int = 256,
                      stream_fn: Callable[[Iterable[torch.Tensor], Iterable[torch.Tensor]], Iterable[torch.Tensor]] = None
                      ) -> List[np.ndarray]:
    """Enhance a list of 1-D float32 waveforms; returns the enhanced waveforms (same lengths).

    ``model`` is a ``wave_u_net_for_speech_enhancement_b200.Model`` on a CUDA device in eval mode. ``stream_fn`` (tests
    only) replaces ``model.forward_host_stream``."""
    frames, index = frame_clips(waveforms, sample_length)
    total = frames.shape[0]
    pinned = frames.is_pinned()
    out = _POOL.get("out", total, sample_length, pinned)
    B = min(batch_frames, total)
    nfull, rem = divmod(total, B)
    batches = [frames[i * B:(i + 1) * B] for i in range(nfull)]
    outs = [out[i * B:(i + 1) * B] for i in range(nfull)]
    tail_in = tail_out = None
    if rem:
        # keep the batch size constant (one plan / workspace): the last batch is filled up with silent frames
        tail_in = _POOL.get("tail_in", B, sample_length, pinned)
        tail_out = _POOL.get("tail_out", B, sample_length, pinned)
        tail_in[:rem] = frames[nfull * B:]
        tail_in[rem:] = 0.0
        batches.append(tail_in)
        outs.append(tail_out)
    fn = stream_fn if stream_fn is not None else model.forward_host_stream
    for _ in fn(batches, outs):
        pass
    if rem:
        out[nfull * B:] = tail_out[:rem]
    return unframe_clips(out, index)


# ---------------------------------------------------------------------------------------------------------------------------
# several GPUs (BASELINE.json configs[3]: batch 1024 sharded over 8 GPUs): one process per GPU, clips dealt to the ranks,
# no collective on the data path (frames are independent in eval mode)
# ---------------------------------------------------------------------------------------------------------------------------
def shard_clips(lengths: Sequence[int], world: int, sample_length: int = 16384) -> List[List[int]]:
    """Deal clips to ``world`` ranks so that every rank gets about the same number of 16384-sample frames: longest clip
    first onto the least loaded rank (ties: lowest rank); every rank's list is returned in ascending clip order.
    Deterministic, so every rank computes the same assignment without communicating."""
    frames = [max(1, -(-int(n) // sample_length)) for n in lengths]
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(frames)), key=lambda i: (-frames[i], i)):
        r = min(range(world), key=lambda r: (load[r], r))
        out[r].append(i)
        load[r] += frames[i]
    return [sorted(ix) for ix in out]


def enhance_waveforms_sharded(model, waveforms: Sequence[np.ndarray], rank: int, world: int, sample_length: int = 16384,
                              batch_frames: int = 256, stream_fn=None, gather: bool = False):
    """This rank's share of :func:`enhance_waveforms` over ``world`` ranks (``enhancement.py:49-74`` sharded by clip).

    Returns ``{clip index: enhanced waveform}`` for the clips of this rank; with ``gather=True`` (needs an initialised
    ``torch.distributed`` group) rank 0 gets the full list in clip order — the only communication, and only of results
    (what a single writer of the output directory, ``enhancement.py:73-74``, needs) — and the other ranks get None."""
    mine = shard_clips([int(np.asarray(w).shape[-1]) for w in waveforms], world, sample_length)[rank]
    local = enhance_waveforms(model, [waveforms[i] for i in mine], sample_length, batch_frames, stream_fn) if mine else []
    result = dict(zip(mine, local))
    if not gather:
        return result
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object(result, parts, dst=0)
    if rank != 0:
        return None
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged[i] for i in range(len(waveforms))]
