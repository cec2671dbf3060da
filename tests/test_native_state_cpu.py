"""Ownership of the native contexts under module copying (nn.DataParallel replicas share ``__dict__`` entries with
the original, trainer/base_trainer.py:26-27; copy.deepcopy after a forward). Driven with a stub library: no GPU needed."""
import copy
import ctypes
import gc

import torch

from wave_u_net_for_speech_enhancement_b200 import _lib, unet_basic
from wave_u_net_for_speech_enhancement_b200.unet_basic import Model


class StubLib:
    def __init__(self):
        self.created, self.destroyed = [], []

    def wunet_create(self, n, ci, device, out):
        handle = 0x1000 + len(self.created)
        self.created.append((handle, device))
        ctypes.cast(out, ctypes.POINTER(ctypes.c_void_p))[0] = handle
        return 0

    def wunet_destroy(self, ctx):
        self.destroyed.append(ctx.value if isinstance(ctx, ctypes.c_void_p) else int(ctx))


def _with_stub(monkeypatch):
    stub = StubLib()
    monkeypatch.setattr(_lib, "load", lambda: stub)
    return stub


def test_replica_shares_the_registry_and_never_destroys_a_context(monkeypatch):
    stub = _with_stub(monkeypatch)
    m = Model(n_layers=2, channels_interval=4)
    ctx0 = m._context(torch.device("cuda", 0))
    assert ctx0.value == 0x1000
    replica = m._replicate_for_data_parallel()              # what torch.nn.parallel.replicate does per device
    assert replica._native is m._native
    assert replica._context(torch.device("cuda", 0)).value == ctx0.value        # found, not re-created
    ctx1 = replica._context(torch.device("cuda", 1))         # a replica on another device creates that device's context once
    assert ctx1.value == 0x1001
    del replica
    gc.collect()
    assert stub.destroyed == []                             # the original still holds live handles
    r2 = m._replicate_for_data_parallel()                   # next forward's replica: same contexts again
    assert r2._context(torch.device("cuda", 1)).value == 0x1001
    assert len(stub.created) == 2
    del r2, m
    gc.collect()
    assert sorted(stub.destroyed) == [0x1000, 0x1001]       # each exactly once, when the last owner of the registry went away


def test_deepcopy_after_a_forward_gets_its_own_registry(monkeypatch):
    stub = _with_stub(monkeypatch)
    m = Model(n_layers=2, channels_interval=4)
    m._context(torch.device("cuda", 0))
    m2 = copy.deepcopy(m)                                   # used to raise: ctypes objects containing pointers cannot be pickled
    assert m2._native is not m._native and m2._native.peek() is None
    assert list(m2.state_dict().keys()) == list(m.state_dict().keys())
    m2._context(torch.device("cuda", 0))
    assert len(stub.created) == 2
    del m2
    gc.collect()
    assert stub.destroyed == [0x1001]


def test_apply_invalidates_packed_weights(monkeypatch):
    _with_stub(monkeypatch)
    m = Model(n_layers=2, channels_interval=4)
    ds = m._state(torch.device("cuda", 0))
    ds.weights_key = ("something",)
    m.float()                                               # any nn.Module._apply (.to/.cpu/.float) drops the packed-weight cache key
    assert ds.weights_key is None
    assert isinstance(unet_basic._NativeState(), unet_basic._NativeState)
