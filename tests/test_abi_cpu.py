"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares; entry
points fail loudly (no CPU fallback) when there is no sm_100 device."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("p2l.h", "p2l_debug.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(p2l_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_every_declared_symbol_is_exported():
    from pix2latent_b200 import _lib, native
    L = _lib.lib()
    decl = declared_symbols()
    assert len(decl) >= 25
    missing = [n for n in decl if not hasattr(L, n)]
    assert not missing, missing
    assert set(native.EXPORTED_SYMBOLS) <= set(decl) | {"p2l_last_error"}


def test_conv_args_struct_matches_header():
    from pix2latent_b200 import _lib
    src = open(os.path.join(ROOT, "include", "p2l_debug.h")).read()
    body = src[src.index("typedef struct p2l_conv_args {"):src.index("} p2l_conv_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for stmt in body.split("{", 1)[1].split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        for part in stmt.split(","):
            fields.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())[0])
    assert fields == [f[0] for f in _lib.ConvArgs._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    from pix2latent_b200 import _lib, native
    L = _lib.lib()
    h = C.c_void_p()
    assert L.p2l_create(0, C.byref(h)) != 0
    assert b"no CUDA device" in L.p2l_last_error()
    with pytest.raises(_lib.P2LError):
        native.context()
    from pix2latent_b200.model import BigGAN
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from pix2latent_b200.model import synth
        cfg = synth.BigGANConfig(output_dim=128, num_classes=16, attention_layer_position=3,
                                 layers=[(True, 4, 4), (True, 4, 4), (True, 4, 4), (False, 4, 4), (True, 4, 2), (True, 2, 1)])
        from pix2latent_b200.model.weights import MissingWeights
        with pytest.raises(MissingWeights):
            BigGAN(config=cfg)              # no checkpoint, no opt-in: refuse to build a meaningless generator
        m = BigGAN(config=cfg, allow_synthetic=True)
        assert m.weights_source == "synthetic"
    with pytest.raises(RuntimeError):
        m(z=torch.zeros(1, 128), c=torch.zeros(1, 128))
    with pytest.raises(RuntimeError):
        m.cuda()
