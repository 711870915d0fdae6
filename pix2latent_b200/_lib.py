"""ctypes binding of the C-ABI library ``libp2l.so`` (include/p2l.h, include/p2l_debug.h).

There is exactly one backend: hand-written sm_100a CUDA. If the library is missing we raise —
no CPU fallback, no alternative dispatch (BASELINE.json north_star).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("P2L_LIB") or os.path.join(_HERE, "libp2l.so")

_lib = None


class P2LError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the extension was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise P2LError(
                "pix2latent_b200: %s not found. Build it with `python -m pix2latent_b200.build` "
                "(or __graft_entry__.build()). There is no CPU fallback." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
        # experiment switches from the environment, e.g. P2L_OPTS="pdl=1,deep=1" (see conv_gemm.h set_option)
        for kv in os.environ.get("P2L_OPTS", "").split(","):
            if "=" in kv:
                k, v = kv.split("=", 1)
                _lib.p2l_debug_set_option(k.strip().encode(), int(v))
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().p2l_last_error()
        raise P2LError("libp2l error %d: %s" % (rc, msg.decode() if msg else "?"))


class ConvArgs(C.Structure):
    """Mirror of ``p2l_conv_args`` (include/p2l_debug.h)."""
    _fields_ = [
        ("A", C.c_void_p),
        ("A_N", C.c_int), ("A_H", C.c_int), ("A_W", C.c_int), ("A_C", C.c_int),
        ("a_c0", C.c_int), ("Cin", C.c_int),
        ("B", C.c_void_p),
        ("Cout", C.c_int), ("B_batch", C.c_int), ("kh", C.c_int), ("kw", C.c_int),
        ("pad_h", C.c_int), ("pad_w", C.c_int),
        ("NI", C.c_int), ("H", C.c_int), ("W", C.c_int), ("BN", C.c_int), ("mode", C.c_int),
        ("alpha", C.c_float),
        ("alpha_ptr", C.c_void_p), ("bias", C.c_void_p),
        ("resid", C.c_void_p), ("resid_C", C.c_int), ("resid_shift", C.c_int),
        ("raw", C.c_void_p), ("raw_C", C.c_int),
        ("raw_f32", C.c_void_p), ("raw_f32_C", C.c_int),
        ("aff_a", C.c_void_p), ("aff_s", C.c_void_p), ("aff_stride", C.c_int), ("relu", C.c_int),
        ("act", C.c_void_p), ("act_C", C.c_int), ("act_up", C.c_int),
        ("act_lo", C.c_void_p), ("img_nchw", C.c_void_p),
        ("saved", C.c_void_p), ("saved_C", C.c_int),
        ("stat0", C.c_void_p), ("stat1", C.c_void_p), ("stat_stride", C.c_int),
        ("addin", C.c_void_p), ("addin_C", C.c_int), ("addin_climit", C.c_int),
        ("addin_pool", C.c_int),
        ("dx", C.c_void_p), ("dx_C", C.c_int),
        ("dx_f32", C.c_void_p), ("dx_f32_C", C.c_int),
        ("rowstat", C.c_void_p), ("rowstat_in", C.c_void_p), ("rowstat_nt", C.c_int),
        ("rowsub", C.c_void_p), ("mulin", C.c_void_p), ("mulin_C", C.c_int),
        ("outT", C.c_void_p), ("outT_c0", C.c_int), ("outT_c1", C.c_int), ("tile_reverse", C.c_int),
    ]


def _declare(L):
    L.p2l_last_error.restype = C.c_char_p
    L.p2l_last_error.argtypes = []
    L.p2l_debug_conv.restype = C.c_int
    L.p2l_debug_conv.argtypes = [C.POINTER(ConvArgs), C.c_void_p]
    L.p2l_debug_set_option.restype = None
    L.p2l_debug_set_option.argtypes = [C.c_char_p, C.c_int]
    L.p2l_debug_get_option.restype = C.c_int
    L.p2l_debug_get_option.argtypes = [C.c_char_p]
    L.p2l_debug_profile_get.restype = C.c_int
    L.p2l_debug_profile_get.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int * 7)]
    # the rest of the API is declared by pix2latent_b200.native (it needs the opaque handles)
    from . import native
    native.declare(L)


def set_option(key, value):
    lib().p2l_debug_set_option(key.encode(), int(value))


def get_option(key):
    return int(lib().p2l_debug_get_option(key.encode()))


def profile_records():
    """[(ms, flops, BN, mode, halo, grid, M, N, K)] of the launches recorded since profile_enable(1)."""
    out, i = [], 0
    while True:
        ms, fl, info = C.c_float(), C.c_double(), (C.c_int * 7)()
        if lib().p2l_debug_profile_get(i, C.byref(ms), C.byref(fl), C.byref(info)) != 0:
            break
        out.append((ms.value, fl.value) + tuple(info))
        i += 1
    return out


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
