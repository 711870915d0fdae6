"""Where the generator / perceptual-net weights come from.

The reference loads REAL checkpoints: ``pytorch_pretrained_biggan.BigGAN.from_pretrained`` (/root/reference
pix2latent/model/biggan.py:26-28), ``torch.load('stylegan2-{car,ffhq}-config-f.pt')['g_ema']`` next to the model file
(pix2latent/model/stylegan2.py:53-85) and the ``lpips`` package's weights (pix2latent/loss_functions.py:131). A model built
without them would invert against a meaningless generator, so that is an ERROR here unless the caller opts in to the
seeded synthetic weights (``allow_synthetic=True`` or ``P2L_ALLOW_SYNTHETIC=1``) — which is what tests and ``bench.py``
do, since the build environment has no network.

Resolution order: explicit ``state_dict`` argument > checkpoint file (argument ``checkpoint`` / environment variable /
the reference's own path) > the third-party package when importable > synthetic (opt-in) > ``MissingWeights``.
"""
import os
import warnings

import torch


class MissingWeights(RuntimeError):
    pass


def synthetic_allowed(flag):
    return bool(flag) or os.environ.get("P2L_ALLOW_SYNTHETIC", "0") not in ("", "0")


def load_checkpoint_file(path):
    """torch.load of a checkpoint file -> flat state dict (rosinality files keep the generator under 'g_ema')."""
    ckpt = torch.load(path, map_location="cpu")
    if isinstance(ckpt, dict) and "g_ema" in ckpt:
        ckpt = ckpt["g_ema"]
    if hasattr(ckpt, "state_dict"):
        ckpt = ckpt.state_dict()
    return {k: v for k, v in ckpt.items() if torch.is_tensor(v)}


def strip_spectral_norm(sd):
    """HF BigGAN checkpoints carry spectral-norm parametrisations (``weight_orig`` + ``weight_u`` / ``weight_v``);
    pix2latent removes them (pix2latent/utils/misc.py:150-157: the trained ``weight_orig`` becomes ``weight``)."""
    out = {}
    for k, v in sd.items():
        if k.endswith(("weight_u", "weight_v")):
            continue
        out[k[:-len("_orig")] if k.endswith("weight_orig") else k] = v
    return out


def first_existing(paths):
    for p in paths:
        if p and os.path.exists(p):
            return p
    return None


def resolve(what, state_dict, candidates, from_package, make_synthetic, allow_synthetic, post=None):
    """Common resolution (see the module docstring). ``candidates``: checkpoint paths to try; ``from_package``: callable
    returning a state dict or None; ``make_synthetic``: callable building the seeded stand-in."""
    if state_dict is not None:
        return state_dict, "state_dict"
    path = first_existing(candidates)
    if path is not None:
        sd = load_checkpoint_file(path)
        return (post(sd) if post else sd), path
    sd = from_package() if from_package is not None else None
    if sd is not None:
        return sd, "package"
    if synthetic_allowed(allow_synthetic):
        warnings.warn("%s: no checkpoint found; using SEEDED SYNTHETIC weights of the same architecture (explicit opt-in). "
                      "Results are meaningless as inversions." % what)
        return make_synthetic(), "synthetic"
    raise MissingWeights(
        "%s: no weights found. Looked for: %s. Pass state_dict=..., or checkpoint=<file>, or install the package the "
        "reference uses; allow_synthetic=True (or P2L_ALLOW_SYNTHETIC=1) builds seeded random-init weights instead "
        "(tests / benchmarks only)." % (what, ", ".join(p for p in candidates if p) or "(no paths)"))
