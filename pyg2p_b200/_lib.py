"""ctypes binding of libg2pgpu.so (C ABI declared in include/g2p.h).

The library is the only compute path of this package: there is no CPU fallback.  If the
shared object is missing, `lib()` raises with the build command; if a call fails, `check`
raises `G2PError` carrying the library's message.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libg2pgpu.so')

OK = 0
ERR_UNSUPPORTED = -5
F64, F32 = 0, 1
MODE_RAW, MODE_NEAREST, MODE_INVDIST = 0, 1, 2
MASK_AS_REFERENCE, MASK_PROPAGATE, MASK_FILL, MASK_PROPAGATE_BITS, MASK_MARKED = 0, 1, 2, 3, 4
SUM_PAIRWISE = 0x100                     # G2P_SUM_PAIRWISE, OR-ed into mask_policy
MASK_MARKER = 0x7ff86d61736b0001          # G2P_MASK_MARKER: the float64 NaN that stands for 'masked' in a source field
FLAG_TIE, FLAG_NEAR_TIE, FLAG_SHORT = 1, 2, 4


class G2PError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'libg2pgpu error {code}: {message}')
        self.code = code


class Rotation(C.Structure):
    _fields_ = [('lat_south_pole_deg', C.c_double), ('lon_south_pole_deg', C.c_double)]


class IndexInfo(C.Structure):
    _fields_ = [('n', C.c_int64), ('cells_per_face_side', C.c_int32), ('n_cells', C.c_int64),
                ('radius', C.c_double), ('mean_spacing', C.c_double), ('bytes', C.c_int64)]


class PlanInfo(C.Structure):
    _fields_ = [('m', C.c_int64), ('n', C.c_int64), ('k', C.c_int32), ('tile_rows', C.c_int32),
                ('tile_cols', C.c_int32), ('n_tiles', C.c_int32), ('max_window', C.c_int32),
                ('max_ranges', C.c_int32), ('mean_window', C.c_double), ('bytes', C.c_int64),
                ('block_rows', C.c_int32), ('block_cols', C.c_int32), ('gather_wavefronts', C.c_double)]


MANIP_MAX_IN = 32
OP_NONE, OP_MUL, OP_ADD, OP_DIV = 0, 1, 2, 3
AGG_PICK, AGG_ACCUM, AGG_AVERAGE, AGG_COPY, AGG_WAVERAGE = 0, 1, 2, 3, 4


class ManipItem(C.Structure):
    _fields_ = [('kind', C.c_int32), ('n_in', C.c_int32), ('inputs', C.c_int32 * MANIP_MAX_IN), ('mask_all', C.c_int32),
                ('edge_mask', C.c_int32), ('unit_time', C.c_double), ('divisor', C.c_double), ('w_edge', C.c_double)]


class ManipDesc(C.Structure):
    _fields_ = [('conv_op1', C.c_int32), ('conv_c1', C.c_double), ('conv_op2', C.c_int32), ('conv_c2', C.c_double),
                ('mv_grib', C.c_double), ('cut_off_negative', C.c_int32), ('n_out', C.c_int32),
                ('items', C.POINTER(ManipItem))]


class Correction(C.Structure):
    _fields_ = [('gem', C.c_void_p), ('dem_term', C.c_void_p), ('valid', C.c_void_p), ('mv', C.c_double),
                ('mode', C.c_int32)]


class WriterPost(C.Structure):
    _fields_ = [('clone_mask', C.c_void_p), ('mv_match', C.c_double), ('use_match', C.c_int32), ('mv_write', C.c_double),
                ('out_f32', C.c_int32)]


_vp, _i64, _i32, _f64 = C.c_void_p, C.c_int64, C.c_int, C.c_double
_rotp = C.POINTER(Rotation)

# name -> (restype, argtypes); must list every symbol include/g2p.h declares
PROTOTYPES = {
    'g2p_version': (C.c_int, []),
    'g2p_last_error': (C.c_char_p, []),
    'g2p_device_count': (C.c_int, []),
    'g2p_set_device': (C.c_int, [_i32]),
    'g2p_launch_count': (C.c_int64, []),
    'g2p_latlon_to_xyz': (C.c_int, [_vp, _vp, _i64, _i32, _f64, _rotp, _vp, _vp]),
    'g2p_index_create': (C.c_int, [_vp, _vp, _i64, _f64, _rotp, C.POINTER(_vp), _vp]),
    'g2p_index_create_xyz': (C.c_int, [_vp, _i64, C.POINTER(_vp), _vp]),
    'g2p_index_destroy': (C.c_int, [_vp]),
    'g2p_index_get_info': (C.c_int, [_vp, C.POINTER(IndexInfo)]),
    'g2p_index_get_xyz': (C.c_int, [_vp, _vp, _vp]),
    'g2p_index_max_nn_dist': (C.c_int, [_vp, C.POINTER(C.c_double), _vp]),
    'g2p_index_max_nn_dist_part': (C.c_int, [_vp, _i64, _i64, C.POINTER(C.c_double), _vp]),
    'g2p_knn': (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _f64, _rotp, _i32, _i32, _f64, _vp, _vp, _vp, _vp, _vp]),
    'g2p_knn_block': (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _f64, _rotp, _i32, _i32, _f64, _vp, _vp, _i64, _i64, _vp, _vp,
                                _vp]),
    'g2p_peer_push': (C.c_int, [_vp, C.POINTER(_vp), _i32, _i64, _i64, _i64, _i32, _vp]),
    'g2p_multicast_push': (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i32, _vp]),
    'g2p_weights': (C.c_int, [_i32, _i32, _i64, _i64, _f64, _i32, _vp, _vp, _vp]),
    'g2p_adw_distance_sum': (C.c_int, [_vp, _i64, _i32, _i32, _i32, C.POINTER(C.c_double), _vp]),
    'g2p_weights_adw': (C.c_int, [_i64, _i32, _i32, _f64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'g2p_table_pack_rec': (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    'g2p_table_unpack_rec': (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    'g2p_table_compact': (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, C.POINTER(C.c_int64), _vp]),
    'g2p_host_pick': (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i32, _i32]),
    'g2p_apply': (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _i32, _i64, _i64, _f64, _i32, _vp, _vp, _vp]),
    'g2p_mark_masked': (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i64, _vp]),
    'g2p_plan_create': (C.c_int, [_vp, _i64, _i32, _i64, _i64, _i64, C.POINTER(_vp), _vp]),
    'g2p_plan_destroy': (C.c_int, [_vp]),
    'g2p_plan_get_info': (C.c_int, [_vp, C.POINTER(PlanInfo)]),
    'g2p_apply_planned': (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _f64, _i32, _vp, _vp, _vp]),
    'g2p_manip': (C.c_int, [C.POINTER(ManipDesc), _vp, _vp, _i64, _i32, _i64, _vp, _i64, _i64, _vp, _vp, _vp]),
    'g2p_time_interp': (C.c_int, [_vp, _vp, _f64, _i64, _vp, _vp]),
    'g2p_correction_prepare': (C.c_int, [_vp, _i32, _f64, _vp, _f64, _f64, _i64, _vp, _vp, _vp]),
    'g2p_correct': (C.c_int, [_vp, _vp, _i64, _i32, _i64, C.POINTER(Correction), _vp]),
    'g2p_apply_planned_corrected': (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _f64, _i32, C.POINTER(Correction),
                                              _vp, _vp, _vp]),
    'g2p_writer_post_apply': (C.c_int, [_vp, _vp, _i64, _i32, _i64, C.POINTER(WriterPost), _vp]),
    'g2p_apply_planned_post': (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _f64, _i32, C.POINTER(Correction),
                                         C.POINTER(WriterPost), _vp, _vp, _vp]),
    'g2p_apply_planned_fused': (C.c_int, [_vp, _vp, C.POINTER(ManipDesc), _vp, _vp, _i32, _i64, _i64, _f64, _i32,
                                          C.POINTER(Correction), _vp, _vp, _vp]),
}

_lib = None


def lib():
    """Load libg2pgpu.so (once).  Raises if it has not been built - never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                f'or `make -C pyg2p_b200/csrc` (nvcc, sm_100a). pyg2p_b200 has no CPU fallback.')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)      # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != OK:
        msg = lib().g2p_last_error()
        raise G2PError(rc, msg.decode('utf-8', 'replace') if msg else '')
    return rc


def rotation_arg(rotation):
    """(lat_sp, lon_sp) or None -> ctypes pointer (or None)."""
    if rotation is None:
        return None
    return C.byref(Rotation(float(rotation[0]), float(rotation[1])))
