"""ctypes binding of libofg.so (include/ofg.h).  No fallback: a missing library is an error."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libofg.so")

OFG_OK = 0
ERR_NAMES = {-1: "OFG_ERR_CUDA", -2: "OFG_ERR_ARG", -3: "OFG_ERR_PARSE", -4: "OFG_ERR_INCONSISTENT", -5: "OFG_ERR_FIT",
             -6: "OFG_ERR_CAPACITY", -7: "OFG_ERR_UNSUPPORTED", -8: "OFG_ERR_MISSING"}
ERR_PARSE, ERR_INCONSISTENT, ERR_FIT, ERR_CAPACITY, ERR_UNSUPPORTED, ERR_MISSING = -3, -4, -5, -6, -7, -8
STAGE_RAW, STAGE_NORM, STAGE_THRESH, STAGE_CONNECT, STAGE_GRAPH = 1, 2, 3, 4, 5
FLAG_BH, FLAG_CONNECT, FLAG_REVERSE, FLAG_TBH, COL_MASK = 0x80000000, 0x40000000, 0x20000000, 0x10000000, 0x0FFFFFFF

# every symbol include/ofg.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = ["ofg_create", "ofg_destroy", "ofg_last_error", "ofg_set_species", "ofg_set_partition", "ofg_add_hits_text", "ofg_add_hits_gz",
           "ofg_clear_hits", "ofg_synth_hits", "ofg_hits_text_size", "ofg_hits_text_copy", "ofg_set_option", "ofg_set_length_factors", "ofg_build",
           "ofg_thresholds_dev", "ofg_thresholds_mark_complete", "ofg_thresholds_copy", "ofg_export_flags",
           "ofg_import_flags", "ofg_counts", "ofg_pair_csr",
           "ofg_pair_fit", "ofg_graph_csr_size", "ofg_graph_csr_copy", "ofg_graph_text_size", "ofg_graph_text_copy",
           "ofg_last_timing", "ofg_kernel_report", "ofg_stream", "ofg_sync", "ofg_get_stat", "ofg_dendro_matrices", "ofg_wait_for", "ofg_export_flags_peer",
           "ofg_import_flags_peer", "ofg_device_alloc", "ofg_device_free", "ofg_ipc_get_handle", "ofg_ipc_open_handle",
           "ofg_ipc_close_handle", "ofg_graph_dev", "ofg_mcl_create", "ofg_mcl_destroy", "ofg_mcl_last_error", "ofg_mcl_set_params",
           "ofg_mcl_set_budgets", "ofg_mcl_set_graph_host", "ofg_mcl_set_graph_text", "ofg_mcl_set_graph_dev", "ofg_mcl_run", "ofg_mcl_step", "ofg_mcl_interpret",
           "ofg_mcl_iterand_copy", "ofg_mcl_result", "ofg_mcl_labels", "ofg_mcl_clusters_text", "ofg_test_log10", "ofg_test_format", "ofg_test_parse_score", "ofg_test_lmfit"]

_lib = None


def load():
    """Load libofg.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libofg.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.ofg_last_error.restype = ctypes.c_char_p
    lib.ofg_last_error.argtypes = [ctypes.c_void_p]
    lib.ofg_stream.restype = ctypes.c_void_p
    lib.ofg_stream.argtypes = [ctypes.c_void_p]
    lib.ofg_destroy.restype = None
    lib.ofg_destroy.argtypes = [ctypes.c_void_p]
    lib.ofg_mcl_last_error.restype = ctypes.c_char_p
    lib.ofg_mcl_last_error.argtypes = [ctypes.c_void_p]
    lib.ofg_mcl_destroy.restype = None
    lib.ofg_mcl_destroy.argtypes = [ctypes.c_void_p]
    vp, i32, i64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t
    sigs = {
        "ofg_create": [ctypes.POINTER(vp), i32],
        "ofg_set_species": [vp, i32, vp, vp],
        "ofg_set_partition": [vp, i32, vp, i32],
        "ofg_add_hits_text": [vp, i32, i32, vp, sz, i32, i32],
        "ofg_add_hits_gz": [vp, i32, i32, vp, sz, i32],
        "ofg_clear_hits": [vp],
        "ofg_synth_hits": [vp, vp, vp],
        "ofg_hits_text_size": [vp, i32, i32, ctypes.POINTER(sz)],
        "ofg_hits_text_copy": [vp, i32, i32, vp, sz],
        "ofg_set_option": [vp, ctypes.c_char_p, i64],
        "ofg_set_length_factors": [vp, vp],
        "ofg_build": [vp, i32],
        "ofg_thresholds_dev": [vp, ctypes.POINTER(vp), ctypes.POINTER(i64)],
        "ofg_thresholds_mark_complete": [vp],
        "ofg_thresholds_copy": [vp, vp, i64],
        "ofg_export_flags": [vp, i32, i32, vp, i32, vp, ctypes.POINTER(vp)],
        "ofg_import_flags": [vp, i32, vp, i64],
        "ofg_sync": [vp],
        "ofg_get_stat": [vp, ctypes.c_char_p, ctypes.POINTER(i64)],
        "ofg_dendro_matrices": [vp, i64, vp, vp, vp, i32, vp, i64, vp],
        "ofg_wait_for": [vp, vp],
        "ofg_export_flags_peer": [vp, i32, i32, vp, i32, vp, sz],
        "ofg_import_flags_peer": [vp, i32, i32, i32, vp, sz],
        "ofg_device_alloc": [vp, sz, ctypes.POINTER(vp)],
        "ofg_device_free": [vp, vp],
        "ofg_ipc_get_handle": [vp, vp, vp],
        "ofg_ipc_open_handle": [vp, vp, ctypes.POINTER(vp)],
        "ofg_ipc_close_handle": [vp, vp],
        "ofg_counts": [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64)],
        "ofg_pair_csr": [vp, i32, i32, ctypes.POINTER(i64), ctypes.POINTER(i64), vp, vp, vp],
        "ofg_pair_fit": [vp, i32, i32, vp],
        "ofg_graph_csr_size": [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)],
        "ofg_graph_csr_copy": [vp, vp, vp, vp, vp],
        "ofg_graph_text_size": [vp, ctypes.POINTER(sz)],
        "ofg_graph_text_copy": [vp, vp, sz],
        "ofg_last_timing": [vp, vp, ctypes.POINTER(i64)],
        "ofg_kernel_report": [vp, ctypes.c_char_p, sz],
        "ofg_graph_dev": [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)],
        "ofg_mcl_create": [ctypes.POINTER(vp), i32],
        "ofg_mcl_set_params": [vp, i32, i32, i32, ctypes.c_double, i32, i32],
        "ofg_mcl_set_budgets": [vp, i64, i64, i32],
        "ofg_mcl_set_graph_host": [vp, i64, vp, vp, vp],
        "ofg_mcl_set_graph_text": [vp, ctypes.c_char_p, sz],
        "ofg_mcl_set_graph_dev": [vp, i64, vp, i64, vp, vp],
        "ofg_mcl_run": [vp, ctypes.c_double],
        "ofg_mcl_step": [vp, ctypes.c_double, ctypes.POINTER(ctypes.c_double)],
        "ofg_mcl_interpret": [vp],
        "ofg_mcl_iterand_copy": [vp, vp, vp, vp],
        "ofg_mcl_result": [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64),
                           ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i64)],
        "ofg_mcl_labels": [vp, vp, i64],
        "ofg_mcl_clusters_text": [vp, vp, sz, ctypes.POINTER(sz)],
        "ofg_test_log10": [vp, vp, vp, i64],
        "ofg_test_format": [vp, vp, vp, i64],
        "ofg_test_parse_score": [vp, ctypes.c_char_p, i64, vp, vp, i64],
        "ofg_test_lmfit": [vp, vp, vp, i64, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = args
    _lib = lib
    return lib
