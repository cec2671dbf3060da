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


def test_frame_clips_padding_is_zero():
    frames, index = enhance.frame_clips([np.ones(S + 3, np.float32)], S, pin=False)
    assert frames.shape == (2, 1, S) and index == [(0, 2, S + 3)]
    assert float(frames[1, 0, 3:].abs().sum()) == 0.0 and float(frames[1, 0, :3].sum()) == 3.0


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
