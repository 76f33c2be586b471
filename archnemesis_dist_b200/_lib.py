"""ctypes binding of libansb200.so (include/ansb200.h).  No torch types cross this boundary:
device pointers and sizes only.  There is no CPU fallback -- if the library is missing the import
fails, and every compute entry point raises on a machine without a CUDA device."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libansb200.so")

OK, EINVAL, ECUDA, ENOMEM = 0, -1, -2, -3
RAD_GRAD, RAD_NAN_TO_NUM, RAD_LAYER_SPACE = 1, 2, 4
MAX_NG, MAX_NGAS = 22, 15
TABLE_F64, TABLE_K32, TABLE_F32 = 0, 1, 2
MAX_LBL_NGAS = 128      # csrc/klbl.cu: gas sum of the line-by-line-table kernel
MAX_NCONV = 65535       # csrc/convolve.cu: grid.y

_vp = ctypes.c_void_p
_i = ctypes.c_int
_d = ctypes.c_double
_u = ctypes.c_uint

EXPORTS = {
    "ansb200_last_error": (ctypes.c_char_p, []),
    "ansb200_version": (_i, []),
    "ansb200_overlap_mode": (_i, [_i]),
    "ansb200_overlap_stats": (None, [ctypes.POINTER(ctypes.c_int32)]),
    "ansb200_table_create": (_i, [_vp, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_vp), _vp]),
    "ansb200_radiance_layer_space": (_i, [_i, _u, _i, _i, _i, _i, _i, _i, _i, _i]),
    "ansb200_jacobian_project_shared": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "ansb200_jacobian_project_chunks": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "ansb200_jacobian_project_sparse": (_i, [_vp] * 5 + [_i, _vp, _vp] + [_i] * 6 + [_vp, _vp]),
    "ansb200_path_mix": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ansb200_kdist_capacity": (_i, [_i]),
    "ansb200_kdist": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "ansb200_table_create_ex": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_vp), _vp]),
    "ansb200_table_destroy": (_i, [_vp]),
    "ansb200_table_shape": (_i, [_vp] + [ctypes.POINTER(_i)] * 5),
    "ansb200_table_k": (_vp, [_vp]),
    "ansb200_table_lnk": (_vp, [_vp]),
    "ansb200_kinterp": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ansb200_koverlap": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ansb200_continuum": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i,
                                _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "ansb200_gas_opacity": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ansb200_radiance": (_i, [_i, _u] + [_vp] * 20 + [_i, _d] + [_i] * 8 + [_vp] * 4),
    "ansb200_jacobian_project": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "ansb200_lbl_absorption": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _d, _d, _d, _d, _d, _d,
                                    _d, _i, _vp, _vp]),
    "ansb200_lbl_table_opacity": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ansb200_convolve": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "ansb200_voigt": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
}

_lib = None


class Ansb200Error(RuntimeError):
    pass


def load():
    """Load libansb200.so (building nothing: see archnemesis_dist_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libansb200.so not found at %s -- run `python -m archnemesis_dist_b200.build` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)     # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        msg = load().ansb200_last_error().decode("utf-8", "replace")
        if rc == EINVAL:
            raise ValueError("ansb200: " + msg)
        if rc == ENOMEM:
            raise MemoryError("ansb200: " + msg)
        raise Ansb200Error("ansb200 (code %d): %s" % (rc, msg))
