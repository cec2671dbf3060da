"""Chunk batching against a literal restatement of the reference loop enhancement.py:49-74."""
import numpy as np
import pytest
import torch

from oracle import wunet_oracle as wo
from wave_u_net_for_speech_enhancement_b200 import enhance

S = 64


def reference_loop(model_fn, clip, sample_length):
    """enhancement.py:53-71 for one clip, one chunk at a time."""
    mixture = torch.from_numpy(clip).reshape(1, 1, -1)
    padded_length = 0
    if mixture.size(-1) % sample_length != 0:
        padded_length = sample_length - (mixture.size(-1) % sample_length)
        mixture = torch.cat([mixture, torch.zeros(1, 1, padded_length)], dim=-1)
    chunks = list(torch.split(mixture, sample_length, dim=-1))
    enhanced = torch.cat([model_fn(c) for c in chunks], dim=-1)
    enhanced = enhanced if padded_length == 0 else enhanced[:, :, :-padded_length]
    return enhanced.reshape(-1).numpy()


def fake_model(frames):                      # frame-wise, position dependent: catches any mis-slicing
    ramp = torch.arange(frames.shape[-1], dtype=torch.float32) / frames.shape[-1]
    return frames * 2.0 + ramp


def fake_stream(batches, outs):
    for b, o in zip(batches, outs):
        o.copy_(fake_model(b))
        yield o


@pytest.mark.parametrize("lengths", [[S], [S * 3], [S * 3 + 5, 1, S - 1, S * 10], [7, 9, 11]])
@pytest.mark.parametrize("batch_frames", [1, 4, 256])
def test_bookkeeping_matches_reference_loop(lengths, batch_frames):
    rng = np.random.default_rng(0)
    clips = [rng.standard_normal(n).astype(np.float32) for n in lengths]
    got = enhance.enhance_waveforms(None, clips, sample_length=S, batch_frames=batch_frames, stream_fn=fake_stream)
    for g, c in zip(got, clips):
        want = reference_loop(fake_model, c, S)
        assert g.shape == want.shape == c.shape
        assert np.array_equal(g, want)
    views = enhance.enhance_waveforms(None, clips, sample_length=S, batch_frames=batch_frames, stream_fn=fake_stream, copy=False)
    assert all(np.array_equal(v, g) for v, g in zip(views, got))          # views of the pooled output buffer, same samples


def test_frame_clips_padding_is_zero():
    frames, index = enhance.frame_clips([np.ones(S + 3, np.float32)], S, pin=False)
    assert frames.shape == (2, 1, S) and index == [(0, 2, S + 3)]
    assert float(frames[1, 0, 3:].abs().sum()) == 0.0 and float(frames[1, 0, :3].sum()) == 3.0


def test_native_framing_ragged_empty_and_reused_buffer():
    """wunet_frame_clips_f32 / wunet_unframe_clips_f32 (row N4) against the per-clip numpy restatement of
    enhancement.py:57-62 / :68-71: ragged lengths, an empty clip (one silent frame), exact multiples, silent frames up to a
    multiple of the batch size, and a staging buffer that still holds the previous call's samples."""
    rng = np.random.default_rng(3)
    enhance.frame_clips([rng.standard_normal(S * 40).astype(np.float32)], S, pin=False)        # dirty the pooled buffer
    lengths = [S * 2 + 7, 0, S, 1, S * 5 - 1, 3 * S]
    clips = [rng.standard_normal(n).astype(np.float32) for n in lengths]
    frames, index = enhance.frame_clips(clips, S, pin=False, round_to=8)
    want_rows = []
    for c in clips:
        nf = max(1, -(-len(c) // S))
        buf = np.zeros(nf * S, np.float32)
        buf[:len(c)] = c
        want_rows.append(buf.reshape(nf, S))
    want = np.concatenate(want_rows)
    assert frames.shape[0] % 8 == 0 and frames.shape[0] >= want.shape[0]
    got = frames.numpy().reshape(frames.shape[0], S)
    assert np.array_equal(got[:want.shape[0]], want)
    assert not got[want.shape[0]:].any()                                                          # silent batch filler
    assert index == [(0, 3, lengths[0]), (3, 1, 0), (4, 1, S), (5, 1, 1), (6, 5, lengths[4]), (11, 3, 3 * S)]
    back = enhance.unframe_clips(frames, index)
    for b, c in zip(back, clips):
        assert b.dtype == np.float32 and np.array_equal(b, c)


def test_native_framing_round_trip_property():
    """Random clip lists (hypothesis): frame -> unframe is the identity, every pad sample is zero, the index matches the
    per-clip chunk counts of enhancement.py:57-62, for any thread count."""
    from hypothesis import given, settings, strategies as hst

    @settings(max_examples=40, deadline=None)
    @given(hst.lists(hst.integers(min_value=0, max_value=5 * S + 3), min_size=1, max_size=9), hst.integers(1, 8), hst.integers(1, 7))
    def check(lengths, threads, round_to):
        rng = np.random.default_rng(sum(lengths) + threads)
        clips = [rng.standard_normal(n).astype(np.float32) for n in lengths]
        old = enhance.HOST_THREADS
        enhance.HOST_THREADS = threads
        try:
            frames, index = enhance.frame_clips(clips, S, pin=False, round_to=round_to)
            flat = frames.numpy().reshape(-1).copy()
            back = enhance.unframe_clips(frames, index)
        finally:
            enhance.HOST_THREADS = old
        assert frames.shape[0] % round_to == 0
        f = 0
        for (f0, nf, n), c, b in zip(index, clips, back):
            assert f0 == f and nf == max(1, -(-n // S)) and n == len(c)
            assert np.array_equal(flat[f0 * S:f0 * S + n], c) and not flat[f0 * S + n:(f0 + nf) * S].any()
            assert np.array_equal(b, c)
            f += nf
        assert not flat[f * S:].any()

    check()


def test_native_framing_int16_pcm_matches_loader_scaling():
    """16-bit PCM clips are converted like librosa.load / soundfile do (sample / 32768), in the framing pass."""
    rng = np.random.default_rng(4)
    pcm = [rng.integers(-32768, 32768, size=n, dtype=np.int16) for n in (S * 2 + 9, S - 1)]
    pcm[0][:3] = [-32768, 32767, 0]
    f16, idx16 = enhance.frame_clips(pcm, S, pin=False)
    f16 = f16.numpy().copy()
    f32, idx32 = enhance.frame_clips([p.astype(np.float32) / 32768.0 for p in pcm], S, pin=False)
    assert idx16 == idx32 and np.array_equal(f16, f32.numpy())
    with pytest.raises(ValueError):
        enhance.frame_clips([pcm[0], pcm[1].astype(np.float32)], S, pin=False)


def test_native_framing_rejects_a_too_small_buffer():
    import ctypes
    from wave_u_net_for_speech_enhancement_b200 import _lib
    lib = _lib.load()
    clip = np.ones(3 * S, np.float32)
    out = np.zeros(2 * S, np.float32)
    ptrs = (ctypes.c_void_p * 1)(clip.ctypes.data)
    lens = (ctypes.c_longlong * 1)(3 * S)
    assert lib.wunet_frame_clips_f32(ptrs, lens, 1, S, out.ctypes.data, 2, 2) < 0
    assert b"frames needed" in lib.wunet_last_error()
    assert not out.any()


@pytest.mark.gpu
def test_gpu_streamed_batches_equal_reference_style_loop():
    """Synthetic 10 s @ 16 kHz clips (BASELINE.json configs[3] shape): batched + streamed == chunk-at-a-time loop."""
    from wave_u_net_for_speech_enhancement_b200 import Model
    st = wo.make_state(12, 24, seed=0)
    m = Model(12, 24, precision="bf16")
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()})
    m = m.to("cuda:0").eval()
    rng = np.random.default_rng(3)
    clips = [(0.3 * rng.standard_normal(n)).astype(np.float32) for n in (160000, 16384, 20000, 163840)]
    got = enhance.enhance_waveforms(m, clips, sample_length=16384, batch_frames=16)

    def model_fn(chunk):                     # enhancement.py:66: model(chunk).detach().cpu()
        return m(chunk.to("cuda:0")).detach().cpu()
    for g, c in zip(got, clips):
        want = reference_loop(model_fn, c, 16384)
        assert g.shape == c.shape
        assert np.array_equal(g, want), "frames are independent: batching must not change a single bit"
