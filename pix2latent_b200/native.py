"""ctypes stub over the opaque-handle C-ABI (include/p2l.h).

This file is the "reference-side binding" of INTEGRATION.md: everything pix2latent's
``closure.step`` needs from the GPU goes through these ~10 C entry points with raw device
pointers and a CUDA stream. torch is used for memory, streams and autograd plumbing only.
"""
import ctypes as C

import torch

from . import _lib

P2L_MAX_LAYERS = 16
LPIPS_NETS = {"alex": 0, "alexnet": 0, "vgg": 1, "vgg16": 1}


class BigGANConfigC(C.Structure):
    """Mirror of ``p2l_biggan_config``."""
    _fields_ = [
        ("n_layers", C.c_int),
        ("up", C.c_int * P2L_MAX_LAYERS), ("in_mult", C.c_int * P2L_MAX_LAYERS),
        ("out_mult", C.c_int * P2L_MAX_LAYERS),
        ("channel_width", C.c_int), ("z_dim", C.c_int), ("class_embed_dim", C.c_int),
        ("attention_pos", C.c_int), ("n_stats", C.c_int),
        ("eps", C.c_float), ("truncation", C.c_float),
    ]


class AdamConfigC(C.Structure):
    """Mirror of ``p2l_adam_config``."""
    _fields_ = [("lr_z", C.c_float), ("lr_c", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("clamp_z", C.c_float), ("clamp_c", C.c_float)]


class SG2ConfigC(C.Structure):
    """Mirror of ``p2l_sg2_config``."""
    _fields_ = [("size", C.c_int), ("style_dim", C.c_int), ("n_mlp", C.c_int), ("channels", C.c_int * 9)]


def declare(L):
    vp, ci, cf, cl = C.c_void_p, C.c_int, C.c_float, C.c_long
    pp = C.POINTER(C.c_void_p)
    sig = {
        "p2l_version": (ci, []),
        "p2l_act_dtype": (ci, []),
        "p2l_launch_count": (cl, []),
        "p2l_create": (ci, [ci, pp]),
        "p2l_destroy": (None, [vp]),
        "p2l_biggan_create": (ci, [vp, C.POINTER(BigGANConfigC), pp]),
        "p2l_biggan_set_tensor": (ci, [vp, C.c_char_p, vp, cl]),
        "p2l_biggan_finalize": (ci, [vp]),
        "p2l_biggan_destroy": (None, [vp]),
        "p2l_biggan_forward": (ci, [vp, ci, vp, vp, vp, vp]),
        "p2l_biggan_backward": (ci, [vp, ci, vp, vp, vp, vp]),
        "p2l_biggan_device_bytes": (cl, [vp]),
        "p2l_biggan_flops": (C.c_double, [vp, ci, ci]),
        "p2l_biggan_launches": (ci, [vp, ci, ci]),
        "p2l_lpips_create": (ci, [vp, ci, pp]),
        "p2l_lpips_set_tensor": (ci, [vp, C.c_char_p, vp, cl]),
        "p2l_lpips_finalize": (ci, [vp]),
        "p2l_lpips_destroy": (None, [vp]),
        "p2l_target_create": (ci, [vp, vp, vp, vp, ci, ci, ci, cf, cf, pp, vp]),
        "p2l_target_destroy": (None, [vp]),
        "p2l_loss_forward": (ci, [vp, vp, ci, vp, vp, ci, vp]),
        "p2l_loss_backward": (ci, [vp, vp, ci, vp, vp, vp]),
        "p2l_lpips_flops": (C.c_double, [vp, ci, ci, ci, ci]),
        "p2l_lpips_launches": (ci, [vp, ci]),
        "p2l_biggan_step": (ci, [vp, vp, vp, ci, vp, vp, ci, cf, vp, vp, vp, vp, vp, vp]),
        "p2l_affine_resample": (ci, [vp, ci, vp, vp, ci, ci, ci, ci, vp]),
        "p2l_biggan_step_targets": (ci, [vp, vp, vp, ci, vp, vp, ci, cf, vp, vp, vp, vp, vp, vp]),
        "p2l_biggan_optimize": (ci, [vp, vp, vp, ci, ci, vp, vp, vp, cf, C.POINTER(AdamConfigC), vp, vp, vp, vp, vp, vp,
                                    ci, vp]),
        "p2l_biggan_optimize_used_graph": (ci, [vp]),
        "p2l_adam_update": (ci, [ci, ci, ci, vp, vp, vp, vp, C.POINTER(AdamConfigC), vp, vp, vp]),
        "p2l_sg2_create": (ci, [vp, C.POINTER(SG2ConfigC), pp]),
        "p2l_sg2_set_tensor": (ci, [vp, C.c_char_p, vp, cl]),
        "p2l_sg2_finalize": (ci, [vp]),
        "p2l_sg2_destroy": (None, [vp]),
        "p2l_sg2_num_noise_layers": (ci, [vp]),
        "p2l_sg2_forward": (ci, [vp, ci, vp, vp, vp, vp]),
        "p2l_sg2_backward": (ci, [vp, ci, vp, vp, vp]),
        "p2l_sg2_step": (ci, [vp, vp, vp, ci, vp, vp, ci, cf, vp, vp, vp, vp, vp]),
        "p2l_sg2_n_latent": (ci, [vp]),
        "p2l_sg2_style": (ci, [vp, ci, vp, vp, vp]),
        "p2l_sg2_forward_w": (ci, [vp, ci, vp, vp, vp, vp]),
        "p2l_sg2_backward_w": (ci, [vp, ci, vp, vp, vp, vp]),
        "p2l_sg2_step_w": (ci, [vp, vp, vp, ci, vp, vp, ci, cf, vp, vp, vp, vp, vp, vp]),
        "p2l_profile_enable": (None, [ci]),
        "p2l_profile_read": (ci, [C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args


EXPORTED_SYMBOLS = [
    "p2l_last_error", "p2l_version", "p2l_act_dtype", "p2l_launch_count", "p2l_create", "p2l_destroy",
    "p2l_biggan_create", "p2l_biggan_set_tensor", "p2l_biggan_finalize", "p2l_biggan_destroy",
    "p2l_biggan_forward", "p2l_biggan_backward", "p2l_biggan_device_bytes", "p2l_biggan_flops",
    "p2l_biggan_launches", "p2l_lpips_create", "p2l_lpips_set_tensor", "p2l_lpips_finalize",
    "p2l_lpips_destroy", "p2l_target_create", "p2l_target_destroy", "p2l_loss_forward",
    "p2l_loss_backward", "p2l_lpips_flops", "p2l_lpips_launches", "p2l_biggan_step",
    "p2l_biggan_optimize", "p2l_biggan_optimize_used_graph", "p2l_adam_update",
    "p2l_affine_resample", "p2l_biggan_step_targets",
    "p2l_sg2_n_latent", "p2l_sg2_style", "p2l_sg2_forward_w", "p2l_sg2_backward_w", "p2l_sg2_step_w",
    "p2l_profile_enable", "p2l_profile_read", "p2l_sg2_create", "p2l_sg2_set_tensor", "p2l_sg2_finalize",
    "p2l_sg2_destroy", "p2l_sg2_num_noise_layers", "p2l_sg2_forward", "p2l_sg2_backward", "p2l_sg2_step",
    "p2l_debug_conv", "p2l_debug_set_option", "p2l_debug_get_option", "p2l_debug_profile_get",
]

_ctx = {}


def _f32c(t):
    assert t.is_cuda, "pix2latent_b200 runs on a CUDA (sm_100a) device only; got a CPU tensor"
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


def context(device=None):
    """One p2l_ctx per CUDA device (created lazily; fails loudly off sm_100)."""
    if not torch.cuda.is_available():
        raise _lib.P2LError("pix2latent_b200 needs a CUDA device (sm_100a); none is visible and "
                            "there is no CPU fallback.")
    dev = torch.cuda.current_device() if device is None else torch.device(device).index or 0
    if dev not in _ctx:
        h = C.c_void_p()
        _lib.check(_lib.lib().p2l_create(dev, C.byref(h)))
        _ctx[dev] = h
    return _ctx[dev], dev


def act_dtype():
    """torch dtype of the library's 16-bit activations / packed weights."""
    return torch.float16 if _lib.lib().p2l_act_dtype() == 1 else torch.bfloat16


def launch_count():
    return int(_lib.lib().p2l_launch_count())


class NativeBigGAN:
    """Owns a ``p2l_biggan`` handle. ``state_dict`` uses pix2latent's BigGAN module keys."""

    def __init__(self, config, state_dict, truncation=1.0, device=None):
        L = _lib.lib()
        self.ctx, self.device = context(device)
        cfg = BigGANConfigC()
        layers = list(config.layers)
        assert len(layers) <= P2L_MAX_LAYERS
        cfg.n_layers = len(layers)
        for i, (up, cin, cout) in enumerate(layers):
            cfg.up[i], cfg.in_mult[i], cfg.out_mult[i] = int(bool(up)), int(cin), int(cout)
        cfg.channel_width = config.channel_width
        cfg.z_dim = config.z_dim
        cfg.class_embed_dim = config.class_embed_dim
        cfg.attention_pos = config.attention_layer_position
        cfg.n_stats = config.n_stats
        cfg.eps = config.eps
        cfg.truncation = truncation
        self.config = config
        self.truncation = truncation
        self.out_res = config.output_dim
        self.h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.p2l_biggan_create(self.ctx, C.byref(cfg), C.byref(self.h)))
            for k, v in state_dict.items():
                if not k.startswith("generator."):
                    continue
                t = v.detach().float().contiguous()
                _lib.check(L.p2l_biggan_set_tensor(self.h, k.encode(), C.c_void_p(t.data_ptr()), t.numel()))
            _lib.check(L.p2l_biggan_finalize(self.h))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().p2l_biggan_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def forward(self, z, c):
        z, c = _f32c(z), _f32c(c)
        b = z.shape[0]
        img = torch.empty(b, 3, self.out_res, self.out_res, device=z.device, dtype=torch.float32)
        _lib.check(_lib.lib().p2l_biggan_forward(self.h, b, _lib.ptr(z), _lib.ptr(c), _lib.ptr(img),
                                                 _lib.current_stream()))
        return img

    def backward(self, b, dimg):
        dimg = _f32c(dimg)
        dz = torch.empty(b, self.config.z_dim, device=dimg.device, dtype=torch.float32)
        dc = torch.empty(b, self.config.class_embed_dim, device=dimg.device, dtype=torch.float32)
        _lib.check(_lib.lib().p2l_biggan_backward(self.h, b, _lib.ptr(dimg), _lib.ptr(dz), _lib.ptr(dc),
                                                  _lib.current_stream()))
        return dz, dc

    def flops(self, b, backward=False):
        return float(_lib.lib().p2l_biggan_flops(self.h, b, int(backward)))

    def device_bytes(self):
        return int(_lib.lib().p2l_biggan_device_bytes(self.h))


class NativeStyleGAN2:
    """Owns a ``p2l_sg2`` handle. ``state_dict`` uses rosinality's ``g_ema`` keys."""

    def __init__(self, size, channels, state_dict, device=None):
        L = _lib.lib()
        self.ctx, self.device = context(device)
        cfg = SG2ConfigC()
        cfg.size, cfg.style_dim, cfg.n_mlp = int(size), 512, 8
        for i in range(9):
            cfg.channels[i] = int(channels.get(2 ** (i + 2), 0))
        self.size = int(size)
        self.h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.p2l_sg2_create(self.ctx, C.byref(cfg), C.byref(self.h)))
            for k, v in state_dict.items():
                if k.startswith("noises.") or k.endswith("blur.kernel") or k.endswith("upsample.kernel"):
                    continue
                t = v.detach().float().contiguous()
                _lib.check(L.p2l_sg2_set_tensor(self.h, k.encode(), C.c_void_p(t.data_ptr()), t.numel()))
            _lib.check(L.p2l_sg2_finalize(self.h))
        self.num_layers = int(L.p2l_sg2_num_noise_layers(self.h))
        self.n_latent = int(L.p2l_sg2_n_latent(self.h))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().p2l_sg2_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def noise_shapes(self, b):
        return [(b, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)) for i in range(self.num_layers)]

    def _noise_ptrs(self, noises, b):
        if noises is None:
            return None, None
        assert len(noises) == self.num_layers, "expected %d noise tensors" % self.num_layers
        keep = [_f32c(n) for n in noises]
        for n, shp in zip(keep, self.noise_shapes(b)):
            assert tuple(n.shape) == shp, "noise shape %s != %s" % (tuple(n.shape), shp)
        arr = (C.c_void_p * self.num_layers)(*[n.data_ptr() for n in keep])
        return arr, keep

    def forward(self, z, noises=None):
        z = _f32c(z)
        b = z.shape[0]
        img = torch.empty(b, 3, self.size, self.size, device=z.device, dtype=torch.float32)
        arr, keep = self._noise_ptrs(noises, b)
        _lib.check(_lib.lib().p2l_sg2_forward(self.h, b, _lib.ptr(z), arr, _lib.ptr(img), _lib.current_stream()))
        return img

    def backward(self, b, dimg):
        dimg = _f32c(dimg)
        dz = torch.empty(b, 512, device=dimg.device, dtype=torch.float32)
        _lib.check(_lib.lib().p2l_sg2_backward(self.h, b, _lib.ptr(dimg), _lib.ptr(dz), _lib.current_stream()))
        return dz


def _sg2_latent(gen, w):
    """[b,512] (w) or [b,n_latent,512] (w+) -> contiguous [b,n_latent,512]."""
    w = _f32c(w)
    if w.dim() == 2:
        w = w.unsqueeze(1).repeat(1, gen.n_latent, 1)
    assert w.dim() == 3 and tuple(w.shape[1:]) == (gen.n_latent, 512), "latent must be [b,512] or [b,%d,512]" % gen.n_latent
    return w.contiguous()


def sg2_style(gen, z):
    """w = style(z): the mapping network alone (stylegan2.py:99-101)."""
    z = _f32c(z)
    w = torch.empty_like(z)
    _lib.check(_lib.lib().p2l_sg2_style(gen.h, z.shape[0], _lib.ptr(z), _lib.ptr(w), _lib.current_stream()))
    return w


def sg2_forward_w(gen, w, noises=None):
    lat = _sg2_latent(gen, w)
    b = lat.shape[0]
    img = torch.empty(b, 3, gen.size, gen.size, device=lat.device, dtype=torch.float32)
    arr, keep = gen._noise_ptrs(noises, b)
    _lib.check(_lib.lib().p2l_sg2_forward_w(gen.h, b, _lib.ptr(lat), arr, _lib.ptr(img), _lib.current_stream()))
    return img


def sg2_backward_w(gen, b, dimg, want_noise_grad=True):
    """(dlatent [b,n_latent,512], [dnoise_l [b,1,r,r]] or None) of the last sg2_forward_w with this b."""
    dimg = _f32c(dimg)
    dlat = torch.empty(b, gen.n_latent, 512, device=dimg.device, dtype=torch.float32)
    dn = [torch.empty(s, device=dimg.device, dtype=torch.float32) for s in gen.noise_shapes(b)] if want_noise_grad else None
    arr = (C.c_void_p * gen.num_layers)(*[t.data_ptr() for t in dn]) if dn is not None else None
    _lib.check(_lib.lib().p2l_sg2_backward_w(gen.h, b, _lib.ptr(dimg), _lib.ptr(dlat), arr, _lib.current_stream()))
    return dlat, dn


def sg2_step_w(gen, lp, tgt, w, noises, want_grad, grad_scale, want_img=True, dloss=None, want_noise_grad=True):
    """Fused StyleGAN2 w/w+ step: returns (loss[b], dlatent [b,n_latent,512], [dnoise_l], img)."""
    lat = _sg2_latent(gen, w)
    b = lat.shape[0]
    dev = lat.device
    loss = torch.empty(b, device=dev, dtype=torch.float32)
    dlat = torch.empty_like(lat) if want_grad else None
    dn = ([torch.empty(s, device=dev, dtype=torch.float32) for s in gen.noise_shapes(b)]
          if (want_grad and want_noise_grad and noises is not None) else None)
    darr = (C.c_void_p * gen.num_layers)(*[t.data_ptr() for t in dn]) if dn is not None else None
    img = torch.empty(b, 3, gen.size, gen.size, device=dev, dtype=torch.float32) if want_img else None
    arr, keep = gen._noise_ptrs(noises, b)
    _lib.check(_lib.lib().p2l_sg2_step_w(gen.h, lp.h, tgt.h, b, _lib.ptr(lat), arr, int(want_grad), float(grad_scale),
                                         _lib.ptr(None if dloss is None else _f32c(dloss)), _lib.ptr(loss), _lib.ptr(dlat),
                                         darr, _lib.ptr(img), _lib.current_stream()))
    return loss, dlat, dn, img


def sg2_step(gen, lp, tgt, z, noises, want_grad, grad_scale, want_img=True, dloss=None):
    """Fused StyleGAN2 inner step: returns (loss[b], dz, img)."""
    z = _f32c(z)
    b = z.shape[0]
    loss = torch.empty(b, device=z.device, dtype=torch.float32)
    dz = torch.empty_like(z) if want_grad else None
    img = torch.empty(b, 3, gen.size, gen.size, device=z.device, dtype=torch.float32) if want_img else None
    arr, keep = gen._noise_ptrs(noises, b)
    _lib.check(_lib.lib().p2l_sg2_step(gen.h, lp.h, tgt.h, b, _lib.ptr(z), arr, int(want_grad), float(grad_scale),
                                       _lib.ptr(None if dloss is None else _f32c(dloss)), _lib.ptr(loss), _lib.ptr(dz),
                                       _lib.ptr(img), _lib.current_stream()))
    return loss, dz, img


class NativeLPIPS:
    def __init__(self, net, state_dict, device=None):
        L = _lib.lib()
        self.ctx, self.device = context(device)
        self.net = net
        self.h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.p2l_lpips_create(self.ctx, LPIPS_NETS[net], C.byref(self.h)))
            for k, v in state_dict.items():
                t = v.detach().float().contiguous()
                _lib.check(L.p2l_lpips_set_tensor(self.h, k.encode(), C.c_void_p(t.data_ptr()), t.numel()))
            _lib.check(L.p2l_lpips_finalize(self.h))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().p2l_lpips_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def make_target(self, target, weight=None, mask=None, rec_type=1, rec_weight=1.0, per_weight=10.0):
        return NativeTarget(self, target, weight, mask, rec_type, rec_weight, per_weight)

    def flops(self, b, H, W, backward=False):
        return float(_lib.lib().p2l_lpips_flops(self.h, b, H, W, int(backward)))


class NativeTarget:
    """Everything that depends only on (target, weight, mask): cached once."""

    def __init__(self, lp, target, weight, mask, rec_type, rec_weight, per_weight):
        self.lp = lp
        t = _f32c(target)
        assert t.dim() == 3 and t.shape[0] == 3, "target must be [3,H,W]"
        self.H, self.W = int(t.shape[1]), int(t.shape[2])
        w = None if weight is None else _f32c(weight.expand_as(t) if weight.shape != t.shape else weight)
        m = None if mask is None else _f32c(mask.expand_as(t) if mask.shape != t.shape else mask)
        self.h = C.c_void_p()
        _lib.check(_lib.lib().p2l_target_create(lp.h, _lib.ptr(t), _lib.ptr(w), _lib.ptr(m), self.H, self.W,
                                                int(rec_type), float(rec_weight), float(per_weight),
                                                C.byref(self.h), _lib.current_stream()))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().p2l_target_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def loss_forward(self, img, want_grad):
        img = _f32c(img)
        b = img.shape[0]
        loss = torch.empty(b, device=img.device, dtype=torch.float32)
        _lib.check(_lib.lib().p2l_loss_forward(self.lp.h, self.h, b, _lib.ptr(img), _lib.ptr(loss),
                                               int(want_grad), _lib.current_stream()))
        return loss

    def loss_backward(self, b, dloss):
        dloss = _f32c(dloss)
        dimg = torch.empty(b, 3, self.H, self.W, device=dloss.device, dtype=torch.float32)
        _lib.check(_lib.lib().p2l_loss_backward(self.lp.h, self.h, b, _lib.ptr(dloss), _lib.ptr(dimg),
                                                _lib.current_stream()))
        return dimg


def profile_enable(on):
    _lib.lib().p2l_profile_enable(int(on))


def profile_read():
    ms, n, fl = C.c_double(), C.c_long(), C.c_double()
    _lib.check(_lib.lib().p2l_profile_read(C.byref(ms), C.byref(n), C.byref(fl)))
    return ms.value, n.value, fl.value


def biggan_step(gen, lp, tgt, z, c, want_grad, grad_scale, want_img=True, dloss=None):
    """One fused inner step (closure.py:51-58) for a mini-batch: returns (loss[b], dz, dc, img).
    Upstream gradient of sample i = grad_scale * (dloss[i] if dloss is given else 1)."""
    z, c = _f32c(z), _f32c(c)
    b = z.shape[0]
    dev = z.device
    loss = torch.empty(b, device=dev, dtype=torch.float32)
    dz = torch.empty_like(z) if want_grad else None
    dc = torch.empty_like(c) if want_grad else None
    img = torch.empty(b, 3, gen.out_res, gen.out_res, device=dev, dtype=torch.float32) if want_img else None
    _lib.check(_lib.lib().p2l_biggan_step(gen.h, lp.h, tgt.h, b, _lib.ptr(z), _lib.ptr(c), int(want_grad),
                                          float(grad_scale), _lib.ptr(None if dloss is None else _f32c(dloss)),
                                          _lib.ptr(loss), _lib.ptr(dz), _lib.ptr(dc),
                                          _lib.ptr(img), _lib.current_stream()))
    return loss, dz, dc, img


def adam_config(lr_z, lr_c, betas=(0.9, 0.999), eps=1e-8, clamp_z=0.0, clamp_c=0.0):
    return AdamConfigC(float(lr_z), float(lr_c), float(betas[0]), float(betas[1]), float(eps),
                       float(clamp_z or 0.0), float(clamp_c or 0.0))


class AdamState:
    """Device-side state of the per-candidate Adam of ``biggan_optimize`` / ``adam_update``:
    ``mv`` [2, b*(zd+cd)] first/second moments laid out (z rows | c rows), ``counters`` int32[2]
    (step count, scratch). Fresh = zeros, like a new torch.optim.Adam (variable_manager.py:238)."""

    def __init__(self, b, zd, cd, device, step=0):
        self.b, self.zd, self.cd = b, zd, cd
        self.mv = torch.zeros(2, b * (zd + cd), device=device, dtype=torch.float32)
        self.counters = torch.tensor([int(step), 0], device=device, dtype=torch.int32)

    def moments(self):
        """(m_z [b,zd], v_z, m_c [b,cd], v_c) views."""
        nz = self.b * self.zd
        m, v = self.mv[0], self.mv[1]
        return (m[:nz].view(self.b, self.zd), v[:nz].view(self.b, self.zd),
                m[nz:].view(self.b, self.cd), v[nz:].view(self.b, self.cd))

    def step_count(self):
        return int(self.counters[0].item())


def adam_update(z, c, dz, dc, cfg, state):
    """In-place Adam update of z [b,zd] and c [b,cd] (closure.py:65 for the latent leaves)."""
    for t in (z, c, dz, dc):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    b = z.shape[0]
    _lib.check(_lib.lib().p2l_adam_update(b, z.shape[1], c.shape[1], _lib.ptr(z), _lib.ptr(c), _lib.ptr(dz), _lib.ptr(dc),
                                          C.byref(cfg), _lib.ptr(state.mv), _lib.ptr(state.counters),
                                          _lib.current_stream()))


def biggan_optimize(gen, lp, tgt, z, c, steps, cfg, state=None, dloss=None, grad_scale=1.0, track=False,
                    want_img=True, use_graph=True):
    """``steps`` fused inner steps (Clamp hook -> generator -> loss -> backward -> Adam) on the device.
    z [b,zd], c [b,cd]: contiguous fp32 CUDA tensors, updated IN PLACE. Returns a dict with
    ``loss`` [steps,b] (per-step losses as closure.step reports them), ``z_hist``/``c_hist`` [steps,b,dim]
    (when ``track``), ``img`` [b,3,R,R] of the last forward, ``state`` (AdamState), ``graph`` (bool)."""
    for t in (z, c):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "z/c must be contiguous fp32 CUDA tensors"
    b, zd, cd = z.shape[0], z.shape[1], c.shape[1]
    dev = z.device
    if state is None:
        state = AdamState(b, zd, cd, dev)
    assert (state.b, state.zd, state.cd) == (b, zd, cd)
    loss = torch.empty(steps, b, device=dev, dtype=torch.float32)
    zh = torch.empty(steps, b, zd, device=dev, dtype=torch.float32) if track else None
    ch = torch.empty(steps, b, cd, device=dev, dtype=torch.float32) if track else None
    img = torch.empty(b, 3, gen.out_res, gen.out_res, device=dev, dtype=torch.float32) if want_img else None
    dl = None if dloss is None else _f32c(dloss)
    _lib.check(_lib.lib().p2l_biggan_optimize(gen.h, lp.h, tgt.h, b, int(steps), _lib.ptr(z), _lib.ptr(c), _lib.ptr(dl),
                                              float(grad_scale), C.byref(cfg), _lib.ptr(state.mv), _lib.ptr(state.counters),
                                              _lib.ptr(loss), _lib.ptr(zh), _lib.ptr(ch), _lib.ptr(img), int(use_graph),
                                              _lib.current_stream()))
    # the library ran on an internal stream ordered before the current one; keep the buffers alive until here
    graph = bool(_lib.lib().p2l_biggan_optimize_used_graph(gen.h))
    return {"loss": loss, "z_hist": zh, "c_hist": ch, "img": img, "state": state, "graph": graph, "_keep": dl}


def affine_resample(src, theta):
    """``F.grid_sample(src, F.affine_grid(theta, size))`` with torch's defaults (bilinear, zero padding,
    align_corners=False) — SpatialTransform.transform (pix2latent/transform/spatial_transform.py:69-85).
    src [b,C,H,W] or [1,C,H,W] (shared source), theta [b,2,3] -> [b,C,H,W]."""
    src, theta = _f32c(src), _f32c(theta)
    assert src.dim() == 4 and theta.dim() == 3 and tuple(theta.shape[1:]) == (2, 3)
    b = theta.shape[0]
    assert src.shape[0] in (1, b), "source batch must be 1 or match theta"
    C_, H, W = int(src.shape[1]), int(src.shape[2]), int(src.shape[3])
    dst = torch.empty(b, C_, H, W, device=src.device, dtype=torch.float32)
    _lib.check(_lib.lib().p2l_affine_resample(_lib.ptr(src), int(src.shape[0]), _lib.ptr(theta), _lib.ptr(dst), b, C_, H, W,
                                              _lib.current_stream()))
    return dst


def biggan_step_targets(gen, lp, tgts, z, c, want_grad, grad_scale, want_img=True, dloss=None):
    """``biggan_step`` with one NativeTarget per candidate (transform search: every candidate's target and
    weight are their own resample). Returns (loss[b], dz, dc, img)."""
    z, c = _f32c(z), _f32c(c)
    b = z.shape[0]
    assert len(tgts) == b, "one target per candidate"
    dev = z.device
    loss = torch.empty(b, device=dev, dtype=torch.float32)
    dz = torch.empty_like(z) if want_grad else None
    dc = torch.empty_like(c) if want_grad else None
    img = torch.empty(b, 3, gen.out_res, gen.out_res, device=dev, dtype=torch.float32) if want_img else None
    arr = (C.c_void_p * b)(*[t.h for t in tgts])
    _lib.check(_lib.lib().p2l_biggan_step_targets(gen.h, lp.h, arr, b, _lib.ptr(z), _lib.ptr(c), int(want_grad),
                                                  float(grad_scale), _lib.ptr(None if dloss is None else _f32c(dloss)),
                                                  _lib.ptr(loss), _lib.ptr(dz), _lib.ptr(dc), _lib.ptr(img),
                                                  _lib.current_stream()))
    return loss, dz, dc, img
