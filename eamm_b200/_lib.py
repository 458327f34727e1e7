"""ctypes binding of libeamm_b200.so (the C ABI declared in include/eamm_b200.h).

There is deliberately no fallback: if the shared object is missing or does not export a symbol the
import fails loudly, and ``check(rc)`` raises on any non-zero return of an entry point.
"""
import ctypes as C
import os

from . import build as _build

EAMM_F32, EAMM_BF16, EAMM_F16 = 0, 1, 2
CONV_3X3, CONV_7X7, CONV_UP2_3X3, CONV_ROW7_PACKED = 0, 1, 2, 3
EPI_RELU, EPI_POOL2, EPI_SIGMOID = 1, 2, 4
SPLITK_WS_BYTES = 4096 + 160 * 128 * 256 * 4          # EAMM_SPLITK_WS_BYTES

_ERR = {-1: "EAMM_ERR_ARG", -2: "EAMM_ERR_SHAPE", -3: "EAMM_ERR_DTYPE", -4: "EAMM_ERR_ALIGN",
        -5: "EAMM_ERR_UNSUPPORTED"}


class Act(C.Structure):
    """eamm_act (include/eamm_b200.h)."""
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("n", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("c", C.c_int32), ("c_off", C.c_int32), ("c_buf", C.c_int32),
                ("planes", C.c_int32), ("n_stride", C.c_int64), ("scale_exp", C.c_int32), ("reserved", C.c_int32)]


class Kp(C.Structure):
    """eamm_kp."""
    _fields_ = [("value", C.c_void_p), ("jacobian", C.c_void_p), ("value_stride", C.c_int64),
                ("jacobian_stride", C.c_int64)]


class OneEuro(C.Structure):
    """eamm_one_euro."""
    _fields_ = [("mincutoff", C.c_float), ("beta", C.c_float), ("dcutoff", C.c_float), ("freq", C.c_float),
                ("scale", C.c_float)]


class ConvArgs(C.Structure):
    """eamm_conv_args."""
    _fields_ = [("kind", C.c_int32), ("flags", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
                ("inp", C.POINTER(Act)), ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("residual", C.POINTER(Act)), ("out", C.POINTER(Act)), ("out2", C.POINTER(Act)),
                ("scale2", C.c_void_p), ("shift2", C.c_void_p), ("out_nchw", C.c_void_p),
                ("out_nchw_c", C.c_int32), ("out_nhwc_f32", C.c_void_p), ("pack_passes", C.c_int32),
                ("weight_fold", C.c_int32), ("out_u8_nhwc", C.c_void_p),
                ("acc_scale", C.c_void_p), ("amax_out", C.c_void_p), ("amax_out2", C.c_void_p),
                ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_int64)]


# name -> (restype, argtypes); must list every symbol of include/eamm_b200.h
_PROTOS = {
    "eamm_abi_version": (C.c_int, []),
    "eamm_device_ok": (C.c_int, [C.c_int]),
    "eamm_aa_downsample": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int, C.c_void_p]),
    "eamm_aa_downsample_act": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_int, C.POINTER(Act), C.c_void_p]),
    "eamm_kp_head": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "eamm_kp_clip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.POINTER(OneEuro), C.POINTER(OneEuro), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "eamm_kp_stage": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(Kp), C.POINTER(Kp), C.c_int, C.c_float,
                                C.POINTER(Act), C.c_void_p, C.c_void_p, C.c_void_p]),
    "eamm_flow_combine": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Kp), C.POINTER(Kp), C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "eamm_warp_occlude": (C.c_int, [C.POINTER(Act), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Act),
                                    C.POINTER(Act), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "eamm_warp_image": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "eamm_nchw_to_act": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Act), C.c_void_p]),
    "eamm_conv_simt": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "eamm_conv_tc": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "eamm_conv_tc_uses_halo": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "eamm_conv_tc_fold": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "eamm_conv_tc_query": (C.c_int, [C.POINTER(ConvArgs), C.POINTER(C.c_int)]),
    "eamm_linear": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                              C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "eamm_maxpool": (C.c_int, [C.POINTER(Act), C.POINTER(Act), C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "eamm_act_copy": (C.c_int, [C.POINTER(Act), C.POINTER(Act), C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "eamm_lstm_layer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "eamm_pack_image": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_PROTOS)


def load():
    """dlopen the in-tree library (building it first if the sources are newer) and bind every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("EAMM_B200_LIB")       # tooling: a diagnostic build of the same sources (tools/experiments/bin)
    if not path:
        path = _build.LIBPATH
        if not os.path.exists(path) or not _build.is_fresh():
            path = _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.eamm_abi_version() != 2:
        raise RuntimeError("libeamm_b200.so ABI version mismatch")
    _lib = lib
    return lib


LAUNCHES = 0   # kernels launched through the C ABI (each entry point launches exactly one)


def check(rc, what=""):
    global LAUNCHES
    if rc == 0:
        LAUNCHES += 1
        return
    if rc < 0:
        raise RuntimeError("eamm_b200: %s rejected its arguments: %s" % (what, _ERR.get(rc, rc)))
    raise RuntimeError("eamm_b200: %s failed with cudaError %d" % (what, rc))
