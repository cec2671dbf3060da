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
