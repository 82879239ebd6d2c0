"""ctypes binding of libscp_b200.so (the C ABI declared in include/scp_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc, and if that is
impossible importing raises.  Calls that need a device raise ``ScpError`` when no sm_100
device is present."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libscp_b200.so")
MAX_DEPTH = 21


class ScpError(RuntimeError):
    pass


class Job(C.Structure):
    _fields_ = [("frame", C.c_int32), ("path_len", C.c_int32), ("path_bits", C.c_int32), ("drop_last", C.c_int32),
                ("qs", C.c_double), ("cart_offset", C.c_double), ("lidar_level", C.c_int32),
                ("pos_eps_last", C.c_int32)]


class JobInfo(C.Structure):
    _fields_ = [("depth", C.c_int32), ("n_points", C.c_int32), ("n_voxels", C.c_int32), ("n_rows", C.c_int32),
                ("row_start", C.c_int64), ("voxel_start", C.c_int64),
                ("level_rows", C.c_int32 * (MAX_DEPTH + 1)), ("bin_num", C.c_float),
                ("steps", C.c_double * 3), ("offset", C.c_double * 3),
                ("pos_min", C.c_int64 * (MAX_DEPTH + 1)), ("pos_max", C.c_int64 * (MAX_DEPTH + 1))]


class OctreeOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("occ", "level", "octant", "parent", "pos", "ctx", "pos_norm", "ctx_pos",
                                           "rows_i64", "voxel_key", "sym")]


class LegacyNode(C.Structure):      # Octreewarpper.py:6-14
    _fields_ = [("nodeid", C.c_uint), ("octant", C.c_uint), ("parent", C.c_uint), ("oct", C.c_uint8),
                ("pos", C.c_uint * 3)]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

SIGNATURES = {
    # name: (restype, argtypes)
    "scp_last_error": (C.c_char_p, []),
    "scp_version": (_i, []),
    "scp_device_ok": (_i, []),
    "scp_launch_count": (_i64, []),
    "scp_octree_create": (_vp, []),
    "scp_octree_destroy": (None, [_vp]),
    "scp_octree_plan": (_i, [_vp, _vp, _i, C.POINTER(_i64), _i, C.POINTER(Job), _i, _i, _vp]),
    "scp_octree_job_info": (_i, [_vp, _i, C.POINTER(JobInfo)]),
    "scp_octree_total_rows": (_i64, [_vp]),
    "scp_octree_total_voxels": (_i64, [_vp]),
    "scp_octree_total_kept": (_i64, [_vp]),
    "scp_set_tree_builder": (_i, [_i]),
    "scp_range_decoder_create": (_vp, [_vp, _i64]),
    "scp_range_decoder_destroy": (None, [_vp]),
    "scp_range_decode": (_i, [_vp, _vp, _i64, _i, _vp]),
    "scp_range_decoder_count": (_i64, [_vp]),
    "scp_decode_level_inputs": (_i, [_vp, _vp, _vp, _i64, _i, _i, C.c_double, C.c_double, _vp, _vp, _vp, _vp]),
    "scp_expand_children": (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp]),
    "scp_dequantise_keys": (_i, [_vp, _i64, _vp, _vp, _i, _vp, _vp]),
    "scp_nn_dist2": (_i, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "scp_octree_emit": (_i, [_vp, C.POINTER(OctreeOut), _vp]),
    "scp_octree_finish": (_i, [_vp, _vp]),
    "scp_octree_stage_ms": (_i, [_vp, C.POINTER(_f * 6)]),
    "scp_segmented_sort_u64": (_i, [_vp, _vp, C.POINTER(_i64), _i, _i, _vp]),
    "new_vector": (_vp, []),
    "delete_vector": (None, [_vp]),
    "vector_size": (_i, [_vp]),
    "vector_get": (_vp, [_vp, _i]),
    "vector_push_back": (None, [_vp, _i]),
    "genOctreeInterface": (_vp, [_vp, C.POINTER(C.c_double), _i]),
    "Nodes_get": (C.POINTER(LegacyNode), [_vp, _i]),
    "Nodes_size": (_i, [_vp]),
    "int_size": (_i, [_vp]),
    "int_get": (_i, [_vp, _i]),
    "scp_coding_order": (_i, [C.POINTER(_i64), _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "scp_gather_windows": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "scp_gather_rows8": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "scp_pad_gather_seqs": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "scp_pmf_to_cdf": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scp_range_encode": (_i64, [_vp, _i64, _vp, _i64]),
    "scp_range_encode_cdf": (_i64, [_vp, _vp, _i64, _i, _vp, _i64]),
    "scp_seqs_create": (_vp, [C.POINTER(_i64), _i]),
    "scp_seqs_create_async": (_vp, [C.POINTER(_i64), _i, _vp]),
    "scp_seqs_destroy": (None, [_vp]),
    "scp_seqs_total": (_i64, [_vp]),
    "scp_linear": (_i, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _i, _i, _vp]),
    "scp_set_auto_engine": (_i, [_i]),
    "scp_set_gemm_cluster": (_i, [_i]),
    "scp_gemm_cache_clear": (None, []),
    "scp_gemm_cache_drop": (None, [_vp]),
    "scp_set_knn_engine": (_i, [_i]),
    "scp_set_attn_engine": (_i, [_i]),
    "scp_linear_tf32_supported": (_i, [_i64, _i64, _i64, _i, _i]),
    "scp_layernorm": (_i, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _i, _f, _vp]),
    "scp_ehem_embed": (_i, [_vp, _i64, _vp, _vp, _i, _vp, _vp, _i64, _vp]),
    "scp_ehem_embed_occ": (_i, [_vp, _i64, _vp, _vp, _i64, _vp]),
    "scp_knn": (_i, [_vp, _i64, _i, _vp, _i, _vp, _vp]),
    "scp_edge_gather_max": (_i, [_vp, _i64, _i, _vp, _i, _i64, _vp, _vp, _vp, _i64, _vp]),
    "scp_edge_gather_max2": (_i, [_vp, _i64, _i, _vp, _i, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "scp_swin_attention": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i64, _vp]),
    "scp_pair_concat": (_i, [_vp, _i64, _vp, _vp, _i, _vp, _i64, _vp]),
    "scp_upsample_cols": (_i, [_vp, _i64, _vp, _vp, _i, _i, _vp, _i64, _i, _vp]),
    "scp_copy_cols": (_i, [_vp, _i64, _i64, _i64, _i64, _i, _vp, _i64, _i, _vp]),
    "scp_octattn_embed": (_i, [_vp, _vp, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scp_octattn_attention": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _i64, _vp]),
    "scp_add": (_i, [_vp, _vp, _vp, _i64, _vp]),
}

_lib = None


def load():
    """Returns the loaded library (building it first if the .so is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if not hasattr(lib, name) and os.environ.get("SCP_DEV_ALLOW_MISSING"):
            continue                      # development only; tests/test_abi_cpu.py checks the full table
        fn = getattr(lib, name)          # AttributeError = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what=""):
    if status is not None and status < 0:
        msg = load().scp_last_error().decode(errors="replace")
        raise ScpError(f"{what} failed ({status}): {msg}")
    return status


_device_ok = {}


def require_device():
    lib = load()
    import torch
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    if _device_ok.get(dev):                 # cudaGetDeviceProperties costs ~20 ms per call: ask once per device
        return lib
    if not lib.scp_device_ok():
        raise ScpError("scp_b200 needs a B200 (sm_100) CUDA device: " + lib.scp_last_error().decode(errors="replace"))
    _device_ok[dev] = True
    return lib


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
