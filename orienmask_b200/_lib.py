"""ctypes binding of the C-ABI library (include/orienmask_b200.h).

There is no fallback: if ``liborienmask_b200.so`` is missing or a call fails, a RuntimeError is
raised.  Build it with ``python -m orienmask_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ORIENMASK_B200_LIB: another build of the same library (A/B of two kernel versions inside one GPU visit)
LIB_PATH = os.environ.get('ORIENMASK_B200_LIB') or os.path.join(HERE, 'liborienmask_b200.so')

OM_MAX_SCALES = 4
OM_MAX_ANCHORS = 16
PREC_F32, PREC_F16, PREC_SPLIT = 0, 1, 2
OUT_ACT, OUT_PARTIAL, OUT_NCHW = 0, 1, 2
NMS_CPU, NMS_CUDA = 0, 1

c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class PostConfig(ctypes.Structure):
    _fields_ = [
        ('num_scales', c_i32), ('num_classes', c_i32), ('image_h', c_i32), ('image_w', c_i32),
        ('grid_h', c_i32 * OM_MAX_SCALES), ('grid_w', c_i32 * OM_MAX_SCALES),
        ('anchors_per_scale', c_i32 * OM_MAX_SCALES), ('anchor_index', (c_i32 * 4) * OM_MAX_SCALES),
        ('total_anchors', c_i32), ('anchor_w', c_f32 * OM_MAX_ANCHORS), ('anchor_h', c_f32 * OM_MAX_ANCHORS),
        ('conf_thresh', c_f32), ('nms_thresh', c_f32), ('orien_thresh', c_f32),
        ('nms_pre', c_i32), ('nms_post', c_i32), ('nms_semantics', c_i32),
    ]


class ConvDesc(ctypes.Structure):
    _fields_ = [
        ('precision', c_i32), ('batch', c_i32),
        ('in_h', c_i32), ('in_w', c_i32), ('in_rows', c_i32),
        ('out_h', c_i32), ('out_w', c_i32), ('out_rows', c_i32),
        ('cin', c_i32), ('cout', c_i32), ('cout_stride', c_i32),
        ('ksize', c_i32), ('stride', c_i32), ('leaky', c_i32), ('out_kind', c_i32),
        ('input', c_vp), ('weights', c_vp), ('bias', c_vp), ('residual', c_vp), ('upadd', c_vp),
        ('up_rows', c_i32), ('output', c_vp), ('in_s2d', c_i32), ('out_s2d', c_i32), ('acc_scale', c_f32),
    ]


class PrepConfig(ctypes.Structure):
    _fields_ = [
        ('src_h', c_i32), ('src_w', c_i32), ('src_dtype', c_i32), ('resize_h', c_i32), ('resize_w', c_i32),
        ('pad_top', c_i32), ('pad_left', c_i32), ('out_h', c_i32), ('out_w', c_i32),
        ('mean', c_f32 * 3), ('std', c_f32 * 3), ('pad_value', c_f32),
    ]


class RleImage(ctypes.Structure):
    _fields_ = [
        ('mask', c_vp), ('count', c_i32), ('mask_h', c_i32), ('mask_w', c_i32), ('top', c_i32), ('left', c_i32),
        ('crop_h', c_i32), ('crop_w', c_i32), ('out_h', c_i32), ('out_w', c_i32), ('hflip', c_i32), ('vflip', c_i32),
    ]


class BlendConfig(ctypes.Structure):
    _fields_ = [('mask_h', c_i32), ('mask_w', c_i32), ('top', c_i32), ('left', c_i32), ('crop_h', c_i32), ('crop_w', c_i32),
                ('out_h', c_i32), ('out_w', c_i32), ('alpha', c_f32)]


class EngineConfig(ctypes.Structure):
    _fields_ = [('precision', c_i32), ('batch', c_i32), ('height', c_i32), ('width', c_i32), ('num_anchors', c_i32),
                ('num_classes', c_i32), ('plus', c_i32)]


class OmTensor(ctypes.Structure):
    _fields_ = [('name', ctypes.c_char_p), ('data', c_vp), ('numel', c_i64)]


class LayerInfo(ctypes.Structure):
    _fields_ = [('name', ctypes.c_char * 64), ('shape', ctypes.c_char * 96), ('flops', ctypes.c_double), ('bytes', ctypes.c_double),
                ('head_slot', c_i32), ('is_stem', c_i32), ('desc', ConvDesc), ('conv', c_vp)]


SRC_U8, SRC_F32 = 0, 1

# name -> (restype, argtypes); every symbol declared in include/orienmask_b200.h
SIGNATURES = {
    'om_abi_version': (c_i32, []),
    'om_last_error': (ctypes.c_char_p, []),
    'om_launch_count': (c_i64, []),
    'om_launch_count_reset': (None, []),
    'om_post_workspace_bytes': (c_i32, [ctypes.POINTER(PostConfig), c_i32, ctypes.POINTER(ctypes.c_size_t)]),
    'om_decode_select': (c_i32, [ctypes.POINTER(PostConfig), ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), c_i32,
                                 c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'om_batched_nms': (c_i32, [ctypes.POINTER(PostConfig), c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'om_mask_assemble': (c_i32, [ctypes.POINTER(PostConfig), ctypes.POINTER(c_vp), ctypes.POINTER(c_i64),
                                 c_vp, c_vp, c_vp, c_i32, c_vp, c_vp]),
    'om_nms': (c_i32, [c_vp, c_i32, c_f32, c_vp, c_vp, c_vp]),
    'om_nms_ex': (c_i32, [c_vp, c_i32, c_f32, c_i32, c_vp, c_vp, c_vp]),
    'om_conv_create': (c_i32, [ctypes.POINTER(ConvDesc), ctypes.POINTER(c_vp)]),
    'om_conv_run': (c_i32, [c_vp, c_vp]),
    'om_conv_run_to': (c_i32, [c_vp, c_vp, c_vp]),
    'om_conv_destroy': (None, [c_vp]),
    'om_preprocess': (c_i32, [ctypes.POINTER(PrepConfig), c_vp, c_i64, c_i32, c_vp, c_vp]),
    'om_mask_rle': (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'om_mask_areas': (c_i32, [ctypes.POINTER(BlendConfig), c_vp, c_i32, c_vp, c_vp, c_vp]),
    'om_mask_blend': (c_i32, [ctypes.POINTER(BlendConfig), c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    'om_stem_conv': (c_i32, [c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    'om_debug_conv_plan_info': (c_i32, [c_vp, ctypes.POINTER(c_i32)]),
    'om_debug_conv_timeline': (c_i32, [c_vp]),
    'om_debug_trace': (c_i32, [c_vp, c_i32]),
    'om_debug_phase_log': (c_i32, [c_vp]),
    'om_engine_workspace_bytes': (c_i32, [ctypes.POINTER(EngineConfig), ctypes.POINTER(ctypes.c_size_t)]),
    'om_engine_create': (c_i32, [ctypes.POINTER(EngineConfig), ctypes.POINTER(OmTensor), c_i32, c_vp, ctypes.c_size_t, c_vp,
                                 ctypes.POINTER(c_vp)]),
    'om_forward': (c_i32, [c_vp, c_vp, ctypes.POINTER(c_vp), c_vp, c_vp]),
    'om_engine_destroy': (None, [c_vp]),
    'om_engine_layer_count': (c_i32, [c_vp]),
    'om_engine_layer_info': (c_i32, [c_vp, c_i32, ctypes.POINTER(LayerInfo)]),
    'om_engine_run_layer': (c_i32, [c_vp, c_i32, c_vp, ctypes.POINTER(c_vp), c_vp, c_vp]),
    'om_engine_run_layers': (c_i32, [c_vp, ctypes.POINTER(c_i32), c_i32, c_vp, ctypes.POINTER(c_vp), c_vp, c_vp]),
}

_lib = None


def lib():
    """The loaded library; raises RuntimeError (never falls back) when it is unavailable."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('orienmask_b200: %s is missing -- run `python -m orienmask_b200.build`; '
                               'there is no CPU or eager fallback' % LIB_PATH)
        try:
            handle = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise RuntimeError('orienmask_b200: cannot load %s: %s' % (LIB_PATH, e))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().om_last_error()
        raise RuntimeError('orienmask_b200.%s failed (%d): %s' % (what, rc, msg.decode() if msg else ''))


def stream_ptr():
    import torch
    return c_vp(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return c_vp(t.data_ptr()) if t is not None else c_vp(0)


def pack_conv_weights(w, precision):
    """[cout, cin, k, k] fp32 (BN already folded) -> (engine layout of om_conv_desc.weights, acc_scale).

    F16: [k*k][cout_pad][cin] half.  F32: [k*k][cin][cout_pad4] float.  SPLIT: [k*k][cout_pad][2*cin] half, hi | lo of the weights
    scaled by a power of two 2^s that puts max|w| in [2^13, 2^14): W_lo = fp16(w*2^s - W_hi) is then a normal fp16 number for
    every weight above 2^-15 of the largest one, and acc_scale = 2^-s undoes the scale in the epilogue (exactly)."""
    import math
    import torch
    cout, cin, k, _ = w.shape
    if precision == PREC_F32:
        cpad = (cout + 3) // 4 * 4
        p = torch.zeros(k * k, cin, cpad, dtype=torch.float32, device=w.device)
        p[:, :, :cout] = w.permute(2, 3, 1, 0).reshape(k * k, cin, cout)
        return p, 1.0
    cpad = (cout + 31) // 32 * 32
    wt = w.float().permute(2, 3, 0, 1).reshape(k * k, cout, cin)
    if precision == PREC_F16:
        p = torch.zeros(k * k, cpad, cin, dtype=torch.float16, device=w.device)
        p[:, :cout] = wt.to(torch.float16)
        return p, 1.0
    amax = float(wt.abs().max())
    s = 13 - math.floor(math.log2(amax)) if amax > 0 else 0
    s = max(-24, min(40, s))
    ws = wt * (2.0 ** s)
    hi = ws.to(torch.float16)
    lo = (ws - hi.float()).to(torch.float16)
    p = torch.zeros(k * k, cpad, 2 * cin, dtype=torch.float16, device=w.device)
    p[:, :cout, :cin] = hi
    p[:, :cout, cin:] = lo
    return p, 2.0 ** (-s)
