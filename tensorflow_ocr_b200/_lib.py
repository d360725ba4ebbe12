"""ctypes binding of libplhead.so (C ABI declared in include/plhead.h).

There is NO fallback: if the shared library has not been built, importing any op
raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C tensorflow_ocr_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("PLH_LIB", "libplhead.so"))   # PLH_LIB: instrumented build for tools/

# ---- constants mirrored from include/plhead.h
OP_LOSS, OP_DECODE, OP_DICE, OP_EAST_LOSS, OP_RESTORE, OP_LOSS_DECODE = range(6)
VARIANT_MODEL, VARIANT_POS_ONLY, VARIANT_PIXELLINK = 0, 1, 2
TERM_CE, TERM_FOCAL = 0, 1
STATS_FLOATS = 64
ST_TOTAL, ST_L_PIX, ST_L_LINK, ST_N_SEG_POS, ST_SUM_WP, ST_SUM_WN = 0, 1, 2, 10, 11, 19
ST_S_PIX, ST_S_POS, ST_S_NEG, ST_LINK_TOTAL, ST_N_SELECTED, ST_THR = 27, 28, 36, 44, 45, 64


class LossParams(C.Structure):
    _fields_ = [("variant", C.c_int32), ("term", C.c_int32), ("neg_pos_ratio", C.c_int32),
                ("focal_alpha", C.c_float), ("focal_gamma", C.c_float), ("reserved", C.c_int32 * 3)]


class DecodeParams(C.Structure):
    _fields_ = [("pixel_thresh", C.c_float), ("link_thresh", C.c_float), ("min_size", C.c_int32),
                ("max_boxes", C.c_int32), ("scale_x", C.c_double), ("scale_y", C.c_double),
                ("reserved", C.c_int32 * 2)]


_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/plhead.h declares
SIGNATURES = {
    "plh_pixellink_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, C.POINTER(LossParams), _vp, _vp, _vp, _vp, _vp,
                                C.POINTER(DecodeParams), _vp, _sz, _vp]),
    "plh_ohnm_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "plh_dice": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _vp, _sz, _vp]),
    "plh_dice_head": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _sz, _vp]),
    "plh_decode": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(DecodeParams), _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "plh_decode_from_flags": (_i, [_vp, _i, _i, _i, C.POINTER(DecodeParams), _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "plh_decode_flags": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(DecodeParams), _vp, _vp]),
    "plh_min_area_boxes": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "plh_pixel_detect": (_i, [_vp, _vp, _i, _i, _f, _f, _vp, _vp]),
    "plh_restore_rectangle": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "plh_restore_rectangle_ex": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "plh_east_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _sz, _vp]),
    "plh_lanms": (_i, [_vp, _vp, _i, _i, C.c_double, _vp, _vp, _vp, _sz, _vp]),
    "plh_link_labels": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "plh_link_labels_icdar": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "plh_contour_workspace_bytes": (_sz, [_i, _i, _i]),
    "plh_contour_boxes": (_i, [_vp, _i, _i, _i, C.c_double, C.c_double, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "plh_head_fuse_level": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plh_fill_quads": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plh_quad_jaccard": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _vp, _vp]),
    "plh_bboxes_matching": (_i, [_vp, _vp, _vp, _vp, _i, _vp, C.c_float, _vp, _vp, _vp, _vp, _vp]),
    "plh_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "plh_version": (_i, []),
    "plh_strerror": (C.c_char_p, [_i]),
    "plh_launch_count": (_ll, []),
    "plh_profile_begin": (_i, [_i]),
    "plh_profile_end": (_i, [C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "plh_profile_kernel_window": (_i, [C.POINTER(C.c_float)]),
}

_lib = None


def load():
    """Load libplhead.so once; raise loudly if it is missing (no CPU / torch fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "tensorflow_ocr_b200: %s is not built. Run `python -c \"import __graft_entry__ as g; g.build()\"` "
            "(needs nvcc; targets sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class PlhError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = load().plh_strerror(rc).decode()
    if rc < 0:
        raise ValueError("%s: %s (plh error %d)" % (what, msg, rc))
    raise PlhError("%s: CUDA error %d: %s" % (what, rc, msg))
