"""ctypes binding of the C ABI declared in include/pyranges_b200.h (libpyranges_b200.so).

The library holds every CUDA kernel of the package; there is no CPU fallback.  If the shared object
is missing the first use raises -- build it with ``make`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYRANGES_B200_LIB") or os.path.join(_HERE, "libpyranges_b200.so")  # (override: kernel A/B builds)

IV_OK = 0
IV_MODE_ALL, IV_MODE_ANY, IV_MODE_CONTAIN, IV_MODE_LAST = 0, 1, 2, 3
IV_JOIN_HINT_ORDERED = 0x100
IV_NEAREST_BOTH, IV_NEAREST_PREVIOUS, IV_NEAREST_NEXT = 0, 1, 2
IV_INDEX_ENDS_PERM = 1
IV_INDEX_GRAPH = 2
IV_ABI_VERSION = 8

_vp, _i64, _i32, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_size_t


class IvGroups(C.Structure):
    _fields_ = [("n_groups", _i32), ("_pad", _i32), ("row_off", _vp), ("lo", _vp), ("hi", _vp), ("base", _vp),
                ("total_span", _i64), ("sum_len", _i64), ("max_len", _i64)]


class IvIndex(C.Structure):
    _fields_ = [("n", _i64), ("total_span", _i64), ("shift", _i32), ("flags", _i32), ("n_bins", _i64),
                ("cross_cap", _i64), ("key_s", _vp), ("row_s", _vp), ("end_s", _vp), ("key_e", _vp), ("row_e", _vp),
                ("soff", _vp), ("seoff", _vp), ("eoff", _vp), ("bins", _vp), ("cross", _vp), ("cross_p", _vp),
                ("groups", IvGroups)]


class IvLists(C.Structure):
    _fields_ = [("n", _i64), ("total_span", _i64), ("shift", _i32), ("reach", _i32), ("n_bins", _i64),
                ("cand_cap", _i64), ("rec", _vp), ("cand_se", _vp), ("cand_row", _vp), ("groups", IvGroups)]


class IvQueries(C.Structure):
    _fields_ = [("n", _i64), ("start", _vp), ("end", _vp), ("label", _vp), ("n_seg", _i32), ("_pad", _i32),
                ("seg_off", _vp), ("seg_group", _vp), ("seg_mode", _vp), ("slack", _i64)]


_PG, _PI, _PQ, _PL = C.POINTER(IvGroups), C.POINTER(IvIndex), C.POINTER(IvQueries), C.POINTER(IvLists)

# name -> (restype, argtypes); mirrors include/pyranges_b200.h one to one
PROTOTYPES = {
    "iv_abi_version": (C.c_int, []),
    "iv_last_error_string": (C.c_char_p, []),
    "iv_set_device": (C.c_int, [C.c_int]),
    "iv_launch_count": (C.c_uint64, []),
    "iv_sort_workspace_bytes": (_sz, [_i64, _i32]),
    "iv_radix_sort_u64": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _vp, _sz, _vp]),
    "iv_segment_bounds": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "iv_lists_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "iv_lists_shift": (_i32, [_i64, _i64, _i64]),
    "iv_lists_build": (C.c_int, [_vp, _vp, _i64, _PG, _i64, _vp, _sz, _PL, _vp]),
    "iv_join_workspace_bytes": (_sz, [_i64]),
    "iv_join_pairs": (C.c_int, [_PL, _PQ, _i32, _vp, _vp, _vp, _i64, _vp, _sz, _vp, _vp]),
    "iv_join_count": (C.c_int, [_PL, _PQ, _i32, _vp, _vp]),
    "iv_sorted_pair_segments": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _vp, _vp]),
    "iv_index_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "iv_index_build": (C.c_int, [_vp, _vp, _i64, _PG, _i32, _vp, _sz, _PI, _vp]),
    "iv_overlap_workspace_bytes": (_sz, [_i64]),
    "iv_count_overlaps": (C.c_int, [_PI, _PQ, _i32, _vp, _vp, _sz, _vp, _vp]),
    "iv_emit_pairs": (C.c_int, [_PI, _PQ, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "iv_segment_pair_counts": (C.c_int, [_PI, _PQ, _vp, _vp, _vp, _vp]),
    "iv_coverage_bp": (C.c_int, [_PI, _PQ, _vp, _vp, _vp]),
    "iv_subtract_workspace_bytes": (_sz, [_i64]),
    "iv_subtract_count": (C.c_int, [_PI, _PQ, _vp, _vp, _sz, _vp, _vp]),
    "iv_subtract_emit": (C.c_int, [_PI, _PQ, _vp, _vp, _vp, _vp, _vp]),
    "iv_nearest": (C.c_int, [_PI, _PQ, _i32, _vp, _vp, _vp]),
    "iv_text_workspace_bytes": (_sz, [_i64]),
    "iv_text_count_lines": (C.c_int, [_vp, _i64, _vp, _sz, _vp, _vp]),
    "iv_text_line_starts": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "iv_bed_parse": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "iv_text_gather": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "iv_k_nearest_workspace_bytes": (_sz, [_i64]),
    "iv_k_nearest_count": (C.c_int, [_PI, _PQ, _vp, _i32, _vp, _vp, _sz, _vp, _vp]),
    "iv_k_nearest_emit": (C.c_int, [_PI, _PQ, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "iv_cluster_workspace_bytes": (_sz, [_i64]),
    "iv_cluster": (C.c_int, [_vp, _vp, _i64, _PG, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "iv_merge_count": (C.c_int, [_vp, _vp, _i64, _PG, _i64, _i32, _vp, _vp, _sz, _vp]),
    "iv_merge_emit": (C.c_int, [_i64, _i64, _PG, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "iv_rle_workspace_bytes": (_sz, [_i64, _i32]),
    "iv_rle_count": (C.c_int, [_vp, _vp, _i64, _PG, _vp, _sz, _vp, _vp]),
    "iv_rle_emit": (C.c_int, [_i64, _PG, _vp, _i64, _vp, _vp, _vp, _vp]),
    "iv_rle_weighted_workspace_bytes": (_sz, [_i64, _i32]),
    "iv_rle_weighted_count": (C.c_int, [_vp, _vp, _vp, _i64, _PG, _vp, _sz, _vp, _vp]),
    "iv_rle_weighted_emit": (C.c_int, [_i64, _PG, _vp, _i64, _vp, _vp, _vp, _vp]),
    "iv_split_workspace_bytes": (_sz, [_i64]),
    "iv_split_count": (C.c_int, [_vp, _vp, _i64, _PG, _vp, _sz, _vp, _vp]),
    "iv_split_emit": (C.c_int, [_i64, _PG, _vp, _i64, _vp, _vp, _vp, _vp]),
    "iv_materialize_pairs": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "iv_gather": (C.c_int, [_vp, _i32, _vp, _i64, _vp, _vp, _vp]),
}

_lib = None


class CabiError(RuntimeError):
    pass


def lib():
    """dlopen libpyranges_b200.so once; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CabiError(
                "pyranges_b200: %s is missing -- the CUDA library has not been built (run `make` in the repo root). "
                "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        v = L.iv_abi_version()
        if v != IV_ABI_VERSION:
            raise CabiError("libpyranges_b200.so ABI version %d != expected %d" % (v, IV_ABI_VERSION))
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != IV_OK:
        msg = lib().iv_last_error_string()
        raise CabiError("%s failed (%d): %s" % (what or "libpyranges_b200 call", rc, msg.decode() if msg else ""))
