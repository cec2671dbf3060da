"""
Training-side data path: the reference's ``dataset/waveform_dataset.py`` with the per-item work done by
the library's host code instead of librosa + numpy slicing.

* :class:`Dataset` is a drop-in for ``dataset.waveform_dataset.Dataset`` (same constructor, same ``(mixture, clean, filename)``
  items, the same ``np.random.randint`` draw per training item as ``util/utils.py:109``, so a seeded run crops the same
  windows) - selected by the ``"module"`` string of the config's dataset stanza, like the model. It reads ONLY the cropped
  window of the two wav files (``wunet_wav_read_f32``) instead of decoding both files completely.
* :class:`CachedPairs` keeps decoded clips (float32, or 16-bit PCM as stored in the wav file) in memory and cuts a whole batch
  of aligned random crops straight into pinned ``[B,1,T]`` batch tensors on several host threads (``wunet_crop_pairs``): the
  pre-framed, pinned input stage for one-process-per-GPU training where the DataLoader workers are the bottleneck.
* :class:`BatchStream` runs that cut for step k+1 on a background thread while step k trains (a ring of pinned buffer pairs).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch.utils import data

from . import _lib

HOST_THREADS = max(1, min(8, (os.cpu_count() or 1) // 2))


def wav_info(path: str) -> dict:
    """Header of a RIFF/WAVE file: sample_rate, channels, frames, bits, is_float."""
    sr, ch, bits, isf = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    fr = ctypes.c_longlong()
    _lib.check(_lib.load().wunet_wav_info(os.fsencode(path), ctypes.byref(sr), ctypes.byref(ch), ctypes.byref(fr),
                                          ctypes.byref(bits), ctypes.byref(isf)))
    return {"sample_rate": sr.value, "channels": ch.value, "frames": fr.value, "bits": bits.value, "is_float": bool(isf.value)}


def load_wav(path: str, start: int = 0, frames: Optional[int] = None) -> Tuple[np.ndarray, int]:
    """``librosa.load(path, sr=None)`` for wav files (dataset/waveform_dataset.py:58-59): mono float32 samples at the file's
    own rate, integer PCM scaled by 2^-(bits-1), channels averaged. ``start`` / ``frames`` select a window of the file."""
    info = wav_info(path)
    n = info["frames"] - start if frames is None else frames
    out = np.empty(max(n, 0), dtype=np.float32)
    _lib.check(_lib.load().wunet_wav_read_f32(os.fsencode(path), start, n, out.ctypes.data_as(ctypes.c_void_p)))
    return out, info["sample_rate"]


class Dataset(data.Dataset):
    """Drop-in for the reference's ``dataset.waveform_dataset.Dataset`` (dataset/waveform_dataset.py:9-67).

    ``dataset``: path of the list file, one ``<noisy path> <clean path>`` pair per line; ``limit`` / ``offset`` select a part
    of the list; ``mode`` "train" returns aligned random crops of ``sample_length`` samples, "validation" the whole signals."""

    def __init__(self, dataset, limit=None, offset=0, sample_length=16384, mode="train"):
        super().__init__()
        dataset_list = [line.rstrip("\n") for line in open(os.path.abspath(os.path.expanduser(dataset)), "r")]
        dataset_list = dataset_list[offset:]
        if limit:
            dataset_list = dataset_list[:limit]
        assert mode in ("train", "validation"), "Mode must be one of 'train' or 'validation'."
        self.length = len(dataset_list)
        self.dataset_list = dataset_list
        self.sample_length = sample_length
        self.mode = mode

    def __len__(self):
        return self.length

    def __getitem__(self, item):
        mixture_path, clean_path = self.dataset_list[item].split(" ")
        filename = os.path.splitext(os.path.basename(mixture_path))[0]
        mixture_path = os.path.abspath(os.path.expanduser(mixture_path))
        clean_path = os.path.abspath(os.path.expanduser(clean_path))
        if self.mode != "train":
            mixture, _ = load_wav(mixture_path)
            clean, _ = load_wav(clean_path)
            return mixture.reshape(1, -1), clean.reshape(1, -1), filename
        # util/utils.py:101-113 (sample_fixed_length_data_aligned): same checks, same draw from numpy's global generator
        na, nb = wav_info(mixture_path)["frames"], wav_info(clean_path)["frames"]
        assert na == nb, "Inconsistent dataset length, unable to sampling"
        assert na >= self.sample_length, f"len(data_a) is {na}, sample_length is {self.sample_length}."
        start = np.random.randint(na - self.sample_length + 1)
        mixture, _ = load_wav(mixture_path, start, self.sample_length)
        clean, _ = load_wav(clean_path, start, self.sample_length)
        return mixture.reshape(1, -1), clean.reshape(1, -1), filename


class CachedPairs:
    """Clips held in memory + batched aligned crops into pinned batch tensors.

    ``pairs``: sequence of ``(mixture, clean)`` - two wav paths, or two 1-D arrays of the same length and dtype (float32, or
    int16 PCM which is converted by /32768 inside the crop like the reference's loader does)."""

    def __init__(self, pairs: Sequence[Tuple], sample_length: int = 16384, pin: bool = True):
        self.sample_length = int(sample_length)
        self.mix: List[np.ndarray] = []
        self.clean: List[np.ndarray] = []
        for a, b in pairs:
            if isinstance(a, (str, os.PathLike)):
                a, _ = load_wav(os.fspath(a))
                b, _ = load_wav(os.fspath(b))
            a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
            if a.dtype != np.int16:
                a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
            assert a.ndim == 1 and b.ndim == 1 and a.dtype == b.dtype
            assert len(a) == len(b), "Inconsistent dataset length, unable to sampling"
            assert len(a) >= self.sample_length, f"len(data_a) is {len(a)}, sample_length is {self.sample_length}."
            self.mix.append(a)
            self.clean.append(b)
        kinds = {a.dtype for a in self.mix}
        assert len(kinds) <= 1, "all clips must share one dtype (float32 or int16)"
        self.is_i16 = kinds == {np.dtype(np.int16)}
        self.pin = bool(pin) and torch.cuda.is_available()
        self._buf = {}

    def __len__(self):
        return len(self.mix)

    def _out(self, name: str, B: int) -> torch.Tensor:
        t = self._buf.get(name)
        if t is None or t.shape[0] < B:
            t = torch.empty(B, 1, self.sample_length, dtype=torch.float32)
            if self.pin:
                t = t.pin_memory()
            self._buf[name] = t
        return t[:B]

    def batch(self, indices: Sequence[int], starts: Optional[Sequence[int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``(mixture, clean)`` batch tensors ``[B,1,T]`` of aligned crops of the clips ``indices``. ``starts`` defaults to one
        ``np.random.randint(len - T + 1)`` per item, in order (util/utils.py:109). The tensors are views of pooled (pinned)
        buffers: valid until the next call."""
        B = len(indices)
        if starts is None:
            starts = [int(np.random.randint(len(self.mix[i]) - self.sample_length + 1)) for i in indices]
        mix = (ctypes.c_void_p * B)(*[self.mix[i].ctypes.data for i in indices])
        cln = (ctypes.c_void_p * B)(*[self.clean[i].ctypes.data for i in indices])
        lens = (ctypes.c_longlong * B)(*[len(self.mix[i]) for i in indices])
        st = (ctypes.c_longlong * B)(*[int(s) for s in starts])
        om, oc = self._out("mixture", B), self._out("clean", B)
        _lib.check(_lib.load().wunet_crop_pairs(mix, cln, lens, st, B, self.sample_length, int(self.is_i16), om.data_ptr(),
                                                oc.data_ptr(), HOST_THREADS))
        return om, oc


class BatchStream:
    """Background producer of training batches over a :class:`CachedPairs`: while the GPU runs step k, ONE host thread cuts the
    crops of step k+1 (``wunet_crop_pairs``, itself multi-threaded) into the next of ``depth`` pinned buffer pairs - what the
    reference gets from DataLoader worker processes (train.py:15-27), without pickling tensors between processes.

    ``batches``: iterable of index lists (one per step; e.g. a shuffled permutation cut into batches). The crop positions are
    drawn in the producer thread, in batch order, from ``rng`` (a ``numpy.random.Generator``; default: a fresh default_rng() -
    NOT numpy's global generator, which is not thread-safe to share with the training thread).
    Iterating yields ``(mixture, clean)`` pinned ``[B,1,T]`` tensors; a pair stays valid until ``depth - 1`` further batches have
    been taken (copy it to the device before that - ``.cuda(non_blocking=True)`` + one stream sync per step is enough)."""

    def __init__(self, pairs: CachedPairs, batches, depth: int = 3, rng: Optional[np.random.Generator] = None):
        import queue
        import threading
        assert depth >= 2
        self.pairs, self.depth = pairs, depth
        self._rng = rng if rng is not None else np.random.default_rng()
        self._q: "queue.Queue" = queue.Queue(maxsize=depth - 1)
        self._free: "queue.Queue" = queue.Queue()
        self._bufs = []
        self._batches = iter(batches)
        self._err = None
        self._thread = threading.Thread(target=self._run, name="wunet-batch-stream", daemon=True)
        self._started = False

    def _buffers(self, B: int):
        T = self.pairs.sample_length
        m = torch.empty(B, 1, T, dtype=torch.float32)
        c = torch.empty(B, 1, T, dtype=torch.float32)
        if self.pairs.pin:
            m, c = m.pin_memory(), c.pin_memory()
        return m, c

    def _run(self):
        try:
            lib = _lib.load()
            for indices in self._batches:
                indices = list(indices)
                B = len(indices)
                slot = self._free.get()                                   # a buffer pair the consumer has released
                if slot is None:
                    return
                m, c = self._bufs[slot]
                if m.shape[0] < B:
                    m, c = self._buffers(B)
                    self._bufs[slot] = (m, c)
                P = self.pairs
                starts = [int(self._rng.integers(len(P.mix[i]) - P.sample_length + 1)) for i in indices]
                mix = (ctypes.c_void_p * B)(*[P.mix[i].ctypes.data for i in indices])
                cln = (ctypes.c_void_p * B)(*[P.clean[i].ctypes.data for i in indices])
                lens = (ctypes.c_longlong * B)(*[len(P.mix[i]) for i in indices])
                st = (ctypes.c_longlong * B)(*starts)
                _lib.check(lib.wunet_crop_pairs(mix, cln, lens, st, B, P.sample_length, int(P.is_i16), m.data_ptr(), c.data_ptr(),
                                                HOST_THREADS))
                self._q.put((slot, m[:B], c[:B], starts))
        except BaseException as e:  # noqa: BLE001 - handed to the consumer
            self._err = e
        finally:
            self._q.put(None)

    def __iter__(self):
        if self._started:
            raise RuntimeError("a BatchStream can be iterated once")
        self._started = True
        for s in range(self.depth):
            self._bufs.append(self._buffers(1))
            self._free.put(s)
        self._thread.start()
        held = []
        try:
            while True:
                item = self._q.get()
                if item is None:
                    if self._err is not None:
                        raise self._err
                    return
                slot, m, c, starts = item
                held.append(slot)
                if len(held) > self.depth - 1:                            # the oldest batch handed out may be overwritten now
                    self._free.put(held.pop(0))
                self.last_starts = starts
                yield m, c
        finally:
            self._free.put(None)                                          # lets a blocked producer exit
