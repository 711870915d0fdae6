"""Weight resolution of the product models (pix2latent_b200/model/weights.py): real checkpoints are required unless the
caller opts in to the seeded synthetic stand-ins (the reference loads torch.load(...)['g_ema'] /
BigGAN.from_pretrained / lpips weights: pix2latent/model/stylegan2.py:83-85, model/biggan.py:26-28,
loss_functions.py:131)."""
import os
import warnings

import pytest
import torch


def test_stylegan2_requires_checkpoint_or_opt_in(tmp_path, monkeypatch):
    from pix2latent_b200.model import synth
    from pix2latent_b200.model.stylegan2 import StyleGAN2
    from pix2latent_b200.model.weights import MissingWeights
    monkeypatch.delenv("P2L_ALLOW_SYNTHETIC", raising=False)
    monkeypatch.delenv("P2L_STYLEGAN2_CKPT", raising=False)
    ch = {4: 64, 8: 64, 16: 64}
    with pytest.raises(MissingWeights):
        StyleGAN2(model="cars", size=16, channels=ch)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = StyleGAN2(model="cars", size=16, channels=ch, allow_synthetic=True)
    assert m.weights_source == "synthetic"
    # a rosinality-style checkpoint file ({'g_ema': state_dict}) under the reference's file name
    sd = synth.stylegan2_state_dict(16, ch, seed=3)
    torch.save({"g_ema": sd}, tmp_path / "stylegan2-car-config-f.pt")
    monkeypatch.setenv("P2L_STYLEGAN2_CKPT", str(tmp_path))
    m = StyleGAN2(model="cars", size=16, channels=ch)
    assert m.weights_source.endswith("stylegan2-car-config-f.pt")
    k = next(iter(sd))
    assert torch.equal(m._state[k].cpu(), sd[k])
    m2 = StyleGAN2(model="cars", size=16, channels=ch, checkpoint=str(tmp_path / "stylegan2-car-config-f.pt"))
    assert torch.equal(m2._state[k].cpu(), sd[k])


def test_biggan_checkpoint_file_with_spectral_norm_keys(tmp_path, monkeypatch):
    from pix2latent_b200.model import BigGAN, synth
    monkeypatch.delenv("P2L_ALLOW_SYNTHETIC", raising=False)
    cfg = synth.BigGANConfig(output_dim=128, num_classes=16, attention_layer_position=3,
                             layers=[(True, 4, 4), (True, 4, 4), (True, 4, 4), (False, 4, 4), (True, 4, 2), (True, 2, 1)])
    sd = synth.biggan_state_dict(cfg, seed=5)
    hf = {}
    for k, v in sd.items():  # as the HF package stores it: spectral-norm parametrised conv / linear weights
        if k.endswith("conv_0.weight"):
            hf[k + "_orig"] = v
            hf[k + "_u"] = torch.zeros(v.shape[0])
            hf[k + "_v"] = torch.zeros(v[0].numel())
        else:
            hf[k] = v
    torch.save(hf, tmp_path / "biggan.pt")
    monkeypatch.setenv("P2L_BIGGAN_CKPT", str(tmp_path / "biggan.pt"))
    m = BigGAN(config=cfg)
    assert m.weights_source.endswith("biggan.pt")
    k = "generator.layers.0.conv_0.weight"
    assert torch.equal(m._state[k].cpu(), sd[k]) and not any(x.endswith("_u") for x in m._state)


def test_lpips_requires_weights_or_opt_in(monkeypatch):
    from pix2latent_b200 import loss_functions as LF
    from pix2latent_b200.model.weights import MissingWeights, resolve
    monkeypatch.delenv("P2L_ALLOW_SYNTHETIC", raising=False)
    monkeypatch.delenv("P2L_LPIPS_CKPT", raising=False)
    with pytest.raises(MissingWeights):
        resolve("LPIPS(alex)", None, [os.environ.get("P2L_LPIPS_CKPT")], lambda: None, lambda: {}, False)
    assert LF.ProjectionLoss()._allow_synthetic is False
    assert LF.ProjectionLoss(allow_synthetic=True)._allow_synthetic is True
    assert LF.ReconstructionLoss()._allow_synthetic is True   # never runs the perceptual net
    monkeypatch.setenv("P2L_ALLOW_SYNTHETIC", "1")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sd, src = resolve("LPIPS(alex)", None, [], lambda: None, lambda: {"x": 1}, False)
    assert src == "synthetic" and sd == {"x": 1}
