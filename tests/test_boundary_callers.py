"""Boundary test (SURVEY §8b recipe): the reference's UNCHANGED callers — `trainer/trainer.py:Trainer`, `trainer/base_trainer.py`
checkpointing, `util/utils.py:initialize_config / load_checkpoint` and the `enhancement.py` script — driven against the drop-in
module selected only by the config's "module" string.

Runs where `/root/reference` exists (the authoring container; it does not travel to the GPU box, where these tests skip).
The packages the reference imports but this image lacks (json5, librosa, pesq, pystoi, matplotlib) are stubbed in
`sys.modules`; nothing of the reference is copied or modified. There is no GPU here, so
  * the training loop runs the drop-in with ``train_backend="torch"`` (the composite that runs on CPU; the native kernels are
    compared with it and with the float64 goldens on the B200 in tests/test_train_gpu.py);
  * for `enhancement.py` the ONE native call (``Model._forward_native``) is replaced by the torch restatement of the same
    formulas — everything around it (plugin loading, strict checkpoint load, `.to()/.eval()`, the [1,1,16384] chunk calls,
    `.detach().cpu()`, trimming) is the real shim. The native forward's numerics are the GPU parity tests' job."""
import copy
import importlib
import json
import os
import runpy
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "trainer")), reason="reference checkout not present on this box")

DROPIN = "wave_u_net_for_speech_enhancement_b200.unet_basic"
N, CI = 4, 8


@pytest.fixture()
def reference_env(monkeypatch):
    """sys.path + stub modules for the reference's third-party imports; undone after the test."""
    wavs = {}                # path -> waveform "on disk"
    written = {}             # path -> waveform written by enhancement.py

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    mod("json5", dumps=lambda o, **k: json.dumps(o, default=str), dump=lambda o, f, **k: json.dump(o, f, default=str),
        load=json.load, loads=json.loads)
    librosa = mod("librosa", load=lambda path, sr=None: (wavs[path].copy(), 16000),
                  stft=lambda y, **k: np.zeros((161, 4), np.complex64), magphase=lambda d: (np.abs(d), d),
                  amplitude_to_db=lambda m: m)
    librosa.output = mod("librosa.output", write_wav=lambda path, y, sr: written.__setitem__(path, np.asarray(y).copy()))
    librosa.display = mod("librosa.display", waveplot=lambda *a, **k: None, specshow=lambda *a, **k: None)
    mod("pesq", pesq=lambda sr, a, b, mode: 2.0)
    pystoi = mod("pystoi")
    pystoi.stoi = mod("pystoi.stoi", stoi=lambda a, b, sr, extended=False: 0.5)
    plt = mod("matplotlib.pyplot", switch_backend=lambda *a: None, subplots=lambda *a, **k: (None, None), tight_layout=lambda: None)
    mod("matplotlib", pyplot=plt)
    monkeypatch.syspath_prepend(REF)
    for name in [k for k in sys.modules if k.split(".")[0] in ("trainer", "util", "model", "dataset")]:
        monkeypatch.delitem(sys.modules, name)             # a fresh import of the reference's packages under the stubs
    yield types.SimpleNamespace(wavs=wavs, written=written)
    for name in [k for k in sys.modules if k.split(".")[0] in ("trainer", "util", "model", "dataset")]:
        sys.modules.pop(name, None)


def train_config(root, epochs):
    return {"seed": 0, "root_dir": str(root), "experiment_name": "boundary", "cudnn_deterministic": False,
            "trainer": {"module": "trainer.trainer", "main": "Trainer", "epochs": epochs, "save_checkpoint_interval": 1,
                        "validation": {"interval": 0, "find_max": True, "custom": {}}},
            "model": {"module": DROPIN, "main": "Model", "args": {"n_layers": N, "channels_interval": CI, "train_backend": "torch"}},
            "loss_function": {"module": "model.loss", "main": "mse_loss", "args": {}},
            "optimizer": {"lr": 0.001, "beta1": 0.9, "beta2": 0.999}}


def synthetic_loader(T=64, batches=3, B=4):
    g = torch.Generator().manual_seed(5)
    out = []
    for i in range(batches):
        clean = 0.1 * torch.randn(B, 1, T, generator=g)
        out.append((clean + 0.05 * torch.randn(B, 1, T, generator=g), clean, [f"f{i}_{j}" for j in range(B)]))
    return out


def build_trainer(cfg, resume):
    """train.py:29-49 with a synthetic loader (the datasets need wav files)"""
    utils = importlib.import_module("util.utils")
    model = utils.initialize_config(cfg["model"])                                   # train.py:29
    optimizer = torch.optim.Adam(params=model.parameters(), lr=cfg["optimizer"]["lr"],
                                 betas=(cfg["optimizer"]["beta1"], cfg["optimizer"]["beta2"]))
    loss_function = utils.initialize_config(cfg["loss_function"])
    trainer_class = utils.initialize_config(cfg["trainer"], pass_args=False)
    return trainer_class(config=cfg, resume=resume, model=model, loss_function=loss_function, optimizer=optimizer,
                         train_dataloader=synthetic_loader(), validation_dataloader=[])


def test_unchanged_trainer_trains_checkpoints_and_resumes_the_dropin(reference_env, tmp_path):
    torch.manual_seed(0)
    t = build_trainer(train_config(tmp_path, 2), resume=False)
    assert type(t.model).__module__ == DROPIN
    before = [p.detach().clone() for p in t.model.parameters()]
    t.train()                                                                       # trainer/base_trainer.py:189-212
    assert any(not torch.equal(a, b) for a, b in zip(before, t.model.parameters())), "optimizer never moved the parameters"
    ck = tmp_path / "boundary" / "checkpoints"
    assert (ck / "latest_model.tar").exists() and (ck / "model_0001.pth").exists() and (ck / "model_0002.pth").exists()

    # the checkpoint the unchanged base_trainer wrote loads strictly into the REFERENCE module and into a fresh drop-in,
    # through the reference's own loader, and both compute the same function
    utils = importlib.import_module("util.utils")
    ref_model_cls = importlib.import_module("model.unet_basic").Model
    sd_pth = utils.load_checkpoint(str(ck / "model_0002.pth"), torch.device("cpu"))
    sd_tar = utils.load_checkpoint(str(ck / "latest_model.tar"), torch.device("cpu"))
    assert list(sd_pth.keys()) == list(sd_tar.keys())
    ref = ref_model_cls(n_layers=N, channels_interval=CI)
    ref.load_state_dict(sd_pth)                                                     # strict
    drop = importlib.import_module(DROPIN).Model(N, CI, train_backend="torch")
    drop.load_state_dict(sd_tar)
    ref.eval(); drop.eval()
    x = torch.randn(2, 1, 64)
    with torch.no_grad():
        assert torch.allclose(ref(x), drop._forward_torch_reference_semantics(x), atol=1e-6)

    # resume (base_trainer.py:62-81): optimizer state (index-keyed) and model state come back, training continues at epoch 3
    t2 = build_trainer(train_config(tmp_path, 3), resume=True)
    assert t2.start_epoch == 3
    for a, b in zip(t2.model.state_dict().values(), sd_tar.values()):
        assert torch.equal(a, b)
    assert len(t2.optimizer.state_dict()["state"]) == len(list(t2.model.parameters()))
    t2.train()
    assert (ck / "model_0003.pth").exists()

    # the nn.DataParallel branches of save / resume (base_trainer.py:76-77, 102-103): `.module` is the drop-in
    t2.model = torch.nn.DataParallel(t2.model)
    t2._save_checkpoint(4)
    t2._resume_checkpoint()
    assert t2.start_epoch == 5 and type(t2.model.module).__module__ == DROPIN


def test_adam_state_written_with_the_reference_model_continues_on_the_dropin(reference_env):
    """Adam's state is keyed by parameter INDEX (base_trainer.py:74,99): a run started with the reference module and resumed
    with the drop-in must continue exactly as the reference would have."""
    ref_model_cls = importlib.import_module("model.unet_basic").Model
    torch.manual_seed(1)
    ref = ref_model_cls(n_layers=N, channels_interval=CI).train()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3, betas=(0.9, 0.999))
    batches = synthetic_loader()
    loss_fn = torch.nn.MSELoss()

    def one_step(model, optimizer, fwd, batch):
        mixture, clean, _ = batch
        optimizer.zero_grad()
        loss = loss_fn(clean, fwd(mixture))
        loss.backward()
        optimizer.step()
        return float(loss.detach())

    one_step(ref, opt, ref, batches[0])
    model_sd = {k: v.clone() for k, v in ref.state_dict().items()}
    opt_sd = copy.deepcopy(opt.state_dict())                                        # state_dict() aliases the live moment tensors
    want = one_step(ref, opt, ref, batches[1])                                      # the reference continuing
    drop = importlib.import_module(DROPIN).Model(N, CI, train_backend="torch").train()
    drop.load_state_dict(model_sd)
    opt2 = torch.optim.Adam(drop.parameters(), lr=1e-3, betas=(0.9, 0.999))
    opt2.load_state_dict(opt_sd)
    got = one_step(drop, opt2, drop, batches[1])
    assert abs(got - want) <= 1e-7 * abs(want)
    for (k, a), b in zip(ref.state_dict().items(), drop.state_dict().values()):
        assert torch.allclose(a.float(), b.float(), atol=1e-7), k


def test_unchanged_enhancement_script_runs_the_dropin(reference_env, tmp_path, monkeypatch):
    """enhancement.py end to end (argparse, config json, dataset, load_checkpoint, chunk loop, write_wav) for the reference
    module and for the drop-in; same checkpoint, same inputs -> same wav files."""
    drop_mod = importlib.import_module(DROPIN)
    monkeypatch.setattr(drop_mod.Model, "_forward_native", lambda self, x: self._forward_torch_reference_semantics(x).detach())
    ref_model_cls = importlib.import_module("model.unet_basic").Model
    torch.manual_seed(2)
    ref = ref_model_cls()                                                           # enhancement config passes "args": {}
    for m in ref.modules():                                                         # eval-BN far from identity
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
    ckpt = tmp_path / "model_0001.pth"
    torch.save(ref.state_dict(), ckpt)
    rng = np.random.default_rng(0)
    listing = tmp_path / "list.txt"
    paths = [str(tmp_path / "a.wav"), str(tmp_path / "b.wav")]
    reference_env.wavs[paths[0]] = (0.3 * rng.standard_normal(16384 + 4000)).astype(np.float32)    # padded to 2 chunks, trimmed
    reference_env.wavs[paths[1]] = (0.3 * rng.standard_normal(16384)).astype(np.float32)           # exactly one chunk
    listing.write_text("\n".join(paths) + "\n")
    results = {}
    for tag, module, main in (("ref", "model.unet_basic", "Model"), ("dropin", DROPIN, "Model")):
        cfg = {"model": {"module": module, "main": main, "args": {}},
               "dataset": {"module": "dataset.waveform_dataset_enhancement", "main": "WaveformDataset",
                           "args": {"dataset": str(listing), "limit": None, "offset": 0, "sample_length": 16384}},
               "custom": {"sample_length": 16384}}
        cfg_path = tmp_path / f"{tag}.json"
        cfg_path.write_text(json.dumps(cfg))
        out_dir = tmp_path / tag
        out_dir.mkdir()
        monkeypatch.setattr(sys, "argv", ["enhancement.py", "-C", str(cfg_path), "-D", "-1", "-O", str(out_dir), "-M", str(ckpt)])
        reference_env.written.clear()
        runpy.run_path(os.path.join(REF, "enhancement.py"), run_name="__main__")
        results[tag] = {os.path.basename(k): v for k, v in reference_env.written.items()}
    assert sorted(results["ref"]) == sorted(results["dropin"]) == ["a.wav", "b.wav"]
    assert results["ref"]["a.wav"].shape == (16384 + 4000,) and results["ref"]["b.wav"].shape == (16384,)
    for k in results["ref"]:
        assert np.abs(results["ref"][k] - results["dropin"][k]).max() <= 1e-6, k


def test_dataset_plugin_is_selected_by_the_config_stanza_and_feeds_the_unchanged_trainer(reference_env, tmp_path):
    """config/train/train.json:37-47 with only the "module" string changed: the reference's own initialize_config builds the
    drop-in Dataset, a stock DataLoader batches it (train.py:15-27), and the unchanged Trainer trains on those batches."""
    import scipy.io.wavfile as wavfile
    from torch.utils.data import DataLoader
    rng = np.random.default_rng(2)
    lines = []
    for i in range(6):
        n = 400 + 13 * i
        clean = rng.integers(-8000, 8000, n, dtype=np.int16)
        noisy = (clean + rng.integers(-900, 900, n)).astype(np.int16)
        wavfile.write(str(tmp_path / f"c{i}.wav"), 16000, clean)
        wavfile.write(str(tmp_path / f"n{i}.wav"), 16000, noisy)
        lines.append(f"{tmp_path / f'n{i}.wav'} {tmp_path / f'c{i}.wav'}")
    (tmp_path / "train.txt").write_text("\n".join(lines) + "\n")
    utils = importlib.import_module("util.utils")
    stanza = {"module": "wave_u_net_for_speech_enhancement_b200.dataset", "main": "Dataset",
              "args": {"dataset": str(tmp_path / "train.txt"), "limit": None, "offset": 0, "sample_length": 64, "mode": "train"}}
    dataset = utils.initialize_config(stanza)                                        # train.py:15
    assert type(dataset).__module__ == "wave_u_net_for_speech_enhancement_b200.dataset" and len(dataset) == 6
    np.random.seed(0)
    loader = DataLoader(dataset=dataset, batch_size=3, num_workers=0, shuffle=False)
    batches = list(loader)
    assert len(batches) == 2 and batches[0][0].shape == (3, 1, 64) and batches[0][0].dtype == torch.float32
    assert list(batches[1][2]) == ["n3", "n4", "n5"]
    cfg = train_config(tmp_path, 1)
    model = utils.initialize_config(cfg["model"])
    optimizer = torch.optim.Adam(params=model.parameters(), lr=1e-3, betas=(0.9, 0.999))
    trainer_class = utils.initialize_config(cfg["trainer"], pass_args=False)
    t = trainer_class(config=cfg, resume=False, model=model, loss_function=utils.initialize_config(cfg["loss_function"]),
                      optimizer=optimizer, train_dataloader=loader, validation_dataloader=[])
    before = [p.detach().clone() for p in t.model.parameters()]
    t.train()
    assert any(not torch.equal(a, b) for a, b in zip(before, t.model.parameters()))
