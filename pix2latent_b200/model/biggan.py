"""BigGAN-deep generator behind pix2latent's model API, executed by the native sm_100a library.

Same surface as /root/reference pix2latent/model/biggan.py:15-58: ``BigGAN(model_version)``,
``.cuda()``, ``.eval()``, ``get_class_embedding(int | onehot[1,1000]) -> [1,128]``,
``forward(z, c, truncation=1.0) -> [b,3,256,256]`` in (-1,1), same asserts. The arithmetic (HF
pytorch_pretrained_biggan Generator) runs in libp2l; backward reaches z and c only — the frozen
generator's weight gradients the reference accumulates are never formed.
"""
import os

import torch
import torch.nn as nn

from .. import native
from . import synth


class _GeneratorFn(torch.autograd.Function):
    """autograd node: image = G(z, c); backward = dgrad to (z, c) through the saved activations."""

    @staticmethod
    def forward(ctx, z, c, model):
        img = model.native.forward(z, c)
        ctx.model = model
        ctx.b = z.shape[0]
        ctx.token = model._bump_forward_token(ctx.b)
        return img

    @staticmethod
    def backward(ctx, dimg):
        model = ctx.model
        if model._forward_token.get(ctx.b) != ctx.token:
            raise RuntimeError(
                "BigGAN.backward: activations of this forward (batch %d) were overwritten by a later "
                "forward with the same batch size; run backward before the next forward" % ctx.b)
        dz, dc = model.native.backward(ctx.b, dimg.contiguous())
        return dz, dc, None


class BigGAN(nn.Module):
    """Drop-in for pix2latent.model.BigGAN.

    Weights (model/weights.py): ``state_dict`` (pix2latent / HF key names) if given; else a checkpoint file
    (``checkpoint=`` or ``$P2L_BIGGAN_CKPT``: a ``torch.save``d HF state dict); else the official checkpoint through
    ``pytorch_pretrained_biggan`` when that package is importable; else — ONLY with ``allow_synthetic=True`` /
    ``P2L_ALLOW_SYNTHETIC=1`` — seeded synthetic weights of the same architecture; otherwise ``MissingWeights``."""

    def __init__(self, model_version="biggan-deep-256", state_dict=None, config=None, seed=0, truncation=1.0,
                 checkpoint=None, allow_synthetic=False):
        super().__init__()
        assert model_version == "biggan-deep-256" or config is not None, \
            "only biggan-deep-256 (or an explicit config) is supported"
        self.config = config or synth.BigGANConfig()
        from . import weights
        state_dict, self.weights_source = weights.resolve(
            "BigGAN(%s)" % model_version, state_dict, [checkpoint, os.environ.get("P2L_BIGGAN_CKPT")],
            (lambda: self._load_pretrained(model_version)) if config is None else None,
            lambda: synth.biggan_state_dict(self.config, seed), allow_synthetic, post=weights.strip_spectral_norm)
        self.register_buffer("embeddings_weight", state_dict["embeddings.weight"].detach().clone().float())
        self._state = {k: v for k, v in state_dict.items() if k.startswith("generator.")}
        self._truncation = float(truncation)
        self.native = None
        self._forward_token = {}
        self._token_counter = 0
        if torch.cuda.is_available():
            self._build()

    @staticmethod
    def _load_pretrained(model_version):
        try:
            import pytorch_pretrained_biggan as ppb
        except ImportError:
            return None
        from ..utils.misc import HiddenPrints, remove_spectral_norm
        with HiddenPrints():
            biggan = ppb.BigGAN.from_pretrained(model_version)
            remove_spectral_norm(biggan.generator)
        return {k: v for k, v in biggan.state_dict().items()}

    def _build(self):
        dev = torch.device("cuda", torch.cuda.current_device())
        sd = {k: v.to(dev) for k, v in self._state.items()}
        self.native = native.NativeBigGAN(self.config, sd, truncation=self._truncation)
        self.embeddings_weight = self.embeddings_weight.to(dev)

    def _bump_forward_token(self, b):
        self._token_counter += 1
        self._forward_token[b] = self._token_counter
        return self._token_counter

    # nn.Module API the reference scripts call
    def cuda(self, device=None):
        if self.native is None:
            if not torch.cuda.is_available():
                raise RuntimeError("BigGAN.cuda(): no CUDA device; pix2latent_b200 has no CPU path")
            self._build()
        return self

    def eval(self):
        return self

    def train(self, mode=True):
        return self

    def get_class_embedding(self, cls):
        """int class label, or a [1, num_classes] one-hot / soft label -> [1, 128] embedding."""
        with torch.no_grad():
            w = self.embeddings_weight
            if type(cls) == int:
                c = torch.zeros(1, self.config.num_classes, device=w.device)
                c[:, cls] = 1
            elif len(cls.size()) == 2:
                c = cls.to(w.device).float()
            else:
                raise ValueError
            return c @ w.t()

    def forward(self, z=None, c=None, truncation=1.0):
        assert 0 < truncation <= 1
        assert len(z.size()) == 2, "expected z to be 2D"
        assert len(c.size()) == 2, "expected c to be 2D"
        assert c.size(1) == self.config.class_embed_dim, \
            "expected c to have dim (?, 128) but got {}".format(c.size())
        if self.native is None:
            raise RuntimeError("BigGAN: native sm_100a model not built (no CUDA device). "
                               "There is no CPU fallback.")
        if abs(truncation - self._truncation) > 1e-12:
            # the BN statistics row is baked in when the weights are packed
            self._truncation = float(truncation)
            self._build()
        if torch.is_grad_enabled() and (z.requires_grad or c.requires_grad):
            return _GeneratorFn.apply(z, c, self)
        return self.native.forward(z, c)
