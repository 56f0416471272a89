"""Build + load liblm_b200.so (the hand-written sm_100a CUDA library) through ctypes.

There is no CPU fallback: if the library cannot be found or built, importing the compute path
raises.  The shared object is built IN-TREE (lagrangian_microbes_b200/liblm_b200.so) so that it
travels with the repository snapshot to the GPU box.
"""
import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "liblm_b200.so")
_SOURCES = ["api.cu", "advect.cu", "bin.cu", "strip.cu", "pairs.cu", "interact.cu", "resolve.cu", "analysis.cu", "record.cu"]
_HEADERS = ["lm_internal.cuh", "philox.cuh", os.path.join("..", "..", "include", "lm_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

LM_OK, LM_EINVAL, LM_ENOMEM, LM_ECUDA, LM_ENOSPC, LM_ESTATE, LM_ENOCONV = 0, -1, -2, -3, -4, -5, -6
LM_STEP_ADVECT, LM_STEP_DIFFUSE, LM_STEP_INTERACT, LM_STEP_EMIT_PAIRS, LM_STEP_STATS = 1, 2, 4, 8, 16
LM_STEP_TIMING = 32
LM_OPT_FIND_PATH, LM_OPT_RESOLVE_UPL, LM_OPT_OVERLAP, LM_OPT_NORM = 2, 3, 4, 5
LM_OPT_RESOLVE_HEAVY_MIN, LM_OPT_RESOLVE_BATCH = 6, 7
LM_OPT_RESOLVE_MODE, LM_OPT_RESOLVE_TILE_SMEM, LM_OPT_RESOLVE_MEGA_MIN, LM_OPT_RESOLVE_TILE_SHAPE = 8, 9, 10, 11
LM_OPT_ADVECT_MODE = 12
LM_OPT_INTERACT_MODE, LM_OPT_DRAW_BATCH, LM_OPT_TILE_CAP = 13, 14, 15
LM_OPT_TILE_REC_CAP, LM_OPT_TILE_PATH, LM_OPT_HEAVY_MIN, LM_OPT_SCATTER_PASSES, LM_OPT_RECORD_DEBUG = 16, 17, 18, 19, 20
LM_OPT_PEER_WAIT_CYCLES = 21
LM_TILE_W, LM_TILE_H = 32, 16
LM_ADVECT_FAITHFUL, LM_ADVECT_FAST = 0, 1
LM_NORM_INF, LM_NORM_1, LM_NORM_2 = 0, 1, 2
LM_PDH_MAX_BINS = 126
LM_FRAME_LAST_DRAWN, LM_FRAME_PLURALITY = 0, 1


class LmError(RuntimeError):
    def __init__(self, code, what, detail=""):
        self.code = code
        super().__init__("%s failed: %s%s" % (what, code, (" -- " + detail) if detail else ""))


class StageTimes(ctypes.Structure):
    _fields_ = [("ti", ctypes.c_int32 * 4), ("interp", ctypes.c_int32 * 4), ("frac", ctypes.c_float * 4)]


class Grid(ctypes.Structure):
    _fields_ = [("x0", ctypes.c_double), ("y0", ctypes.c_double), ("inv_h", ctypes.c_double),
                ("ncx", ctypes.c_int32), ("ncy", ctypes.c_int32)]

    def as_dict(self):
        return dict(x0=self.x0, y0=self.y0, inv_h=self.inv_h, ncx=self.ncx, ncy=self.ncy)


class RpsParams(ctypes.Structure):
    _fields_ = [("pRS", ctypes.c_double), ("pPR", ctypes.c_double), ("pSP", ctypes.c_double),
                ("seed", ctypes.c_uint64), ("step", ctypes.c_uint64)]


class Stats(ctypes.Structure):
    _fields_ = [("n_pairs", ctypes.c_int64), ("n_out_of_bounds", ctypes.c_int64), ("n_clamped", ctypes.c_int64),
                ("species_count", ctypes.c_int64 * 4), ("bbox", ctypes.c_float * 4),
                ("n_particles", ctypes.c_int64), ("n_moved_in", ctypes.c_int64), ("n_moved_out", ctypes.c_int64),
                ("n_misrouted", ctypes.c_int64)]


class Strip(ctypes.Structure):
    _fields_ = [("row0", ctypes.c_int32), ("rows_owned", ctypes.c_int32), ("has_south", ctypes.c_int32),
                ("has_north", ctypes.c_int32)]


LM_IPC_HANDLE_BYTES, LM_PEER_BUFFERS = 64, 6
LM_XCHG_MIG, LM_XCHG_GHOST, LM_XCHG_GSP, LM_XCHG_GRET = 0, 1, 2, 3


class PeerExport(ctypes.Structure):
    """lm_peer_export: IPC handles (another process) and plain device pointers (same process) of a strip's receive buffers + flags."""
    _fields_ = [("ipc", (ctypes.c_ubyte * LM_IPC_HANDLE_BYTES) * LM_PEER_BUFFERS), ("ptr", ctypes.c_void_p * LM_PEER_BUFFERS)]


class StripBuffers(ctypes.Structure):
    _fields_ = [("mig_send", ctypes.c_void_p * 2), ("mig_recv", ctypes.c_void_p * 2), ("mig_bytes", ctypes.c_int64),
                ("ghost_send", ctypes.c_void_p), ("ghost_recv", ctypes.c_void_p), ("ghost_bytes", ctypes.c_int64),
                ("gsp_send", ctypes.c_void_p), ("gsp_recv", ctypes.c_void_p), ("gret_send", ctypes.c_void_p),
                ("gret_recv", ctypes.c_void_p), ("species_bytes", ctypes.c_int64)]


def _stale():
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    files = [os.path.join(_CSRC, f) for f in _SOURCES + _HEADERS]
    return any(os.path.getmtime(f) > t for f in files if os.path.exists(f))


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into liblm_b200.so (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return _SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and %s is missing/stale: cannot build the CUDA library" % _SO)
    tmp = _SO + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + _SOURCES
    subprocess.check_call(cmd, cwd=_CSRC)
    os.replace(tmp, _SO)
    return _SO


_lib = None


def lib():
    """The loaded library with argtypes/restypes declared.  Raises if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    elif _stale() and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        build()
    L = ctypes.CDLL(_SO)
    declare(L)
    _lib = L
    return L


def declare(L):
    """argtypes / restypes of every entry point of include/lm_b200.h on a loaded library."""
    vp, i32, i64, u64, dbl, flt = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64,
                                  ctypes.c_double, ctypes.c_float)
    P = ctypes.POINTER
    sig = {
        "lm_version": (ctypes.c_int, []),
        "lm_error_string": (ctypes.c_char_p, [ctypes.c_int]),
        "lm_last_cuda_error": (ctypes.c_char_p, []),
        "lm_create": (ctypes.c_int, [P(vp), ctypes.c_int, i64, i64, i64]),
        "lm_destroy": (ctypes.c_int, [vp]),
        "lm_set_field": (ctypes.c_int, [vp, vp, vp, vp, vp, i32, i32, i32]),
        "lm_update_field_data": (ctypes.c_int, [vp, vp, vp, i32]),
        "lm_set_grid": (ctypes.c_int, [vp, P(Grid)]),
        "lm_get_grid": (ctypes.c_int, [vp, P(Grid)]),
        "lm_advect_rk4": (ctypes.c_int, [vp, vp, vp, i64, P(StageTimes), flt, vp]),
        "lm_diffuse": (ctypes.c_int, [vp, vp, vp, i64, dbl, u64, u64, vp]),
        "lm_diffuse_ids": (ctypes.c_int, [vp, vp, vp, vp, i64, dbl, u64, u64, vp]),
        "lm_find_pairs": (ctypes.c_int, [vp, vp, vp, i64, dbl, vp, i64, vp, vp]),
        "lm_interact_rps": (ctypes.c_int, [vp, vp, vp, vp, i64, dbl, P(RpsParams), vp, i64, vp, vp]),
        "lm_pair_uniforms": (ctypes.c_int, [vp, i64, u64, u64, vp, vp]),
        "lm_resolve_rps": (ctypes.c_int, [vp, vp, vp, i64, vp, i64, dbl, dbl, dbl, P(i32), vp]),
        "lm_state_set": (ctypes.c_int, [vp, vp, vp, vp, vp, i64, vp]),
        "lm_state_size": (i64, [vp]),
        "lm_step": (ctypes.c_int, [vp, i32, P(StageTimes), flt, dbl, dbl, P(RpsParams), vp, i64, vp]),
        "lm_strip_alloc": (ctypes.c_int, [vp, i64, i64, i32]),
        "lm_set_strip": (ctypes.c_int, [vp, P(Strip)]),
        "lm_strip_buffers_get": (ctypes.c_int, [vp, P(StripBuffers)]),
        "lm_strip_peer_export": (ctypes.c_int, [vp, P(PeerExport)]),
        "lm_strip_peer_connect": (ctypes.c_int, [vp, i32, P(PeerExport), i32]),
        "lm_step_push": (ctypes.c_int, [vp, i32, vp]),
        "lm_step_move": (ctypes.c_int, [vp, i32, P(StageTimes), flt, dbl, P(RpsParams), vp]),
        "lm_step_bin": (ctypes.c_int, [vp, vp]),
        "lm_step_interact_begin": (ctypes.c_int, [vp, dbl, vp, i64, vp]),
        "lm_step_interact_end": (ctypes.c_int, [vp, vp]),
        "lm_step_finish": (ctypes.c_int, [vp, vp]),
        "lm_state_get": (ctypes.c_int, [vp, vp, vp, vp, vp]),
        "lm_state_get_host": (ctypes.c_int, [vp, vp, vp, vp, vp]),
        "lm_record_next_step": (ctypes.c_int, [vp, vp, vp, vp]),
        "lm_record_next_step_ids": (ctypes.c_int, [vp, vp, vp, vp, vp]),
        "lm_record_count": (ctypes.c_int64, [vp]),
        "lm_host_copies_sync": (ctypes.c_int, [vp]),
        "lm_state_view": (ctypes.c_int, [vp, P(vp), P(vp), P(vp), P(vp), P(vp)]),
        "lm_sync_stats": (ctypes.c_int, [vp, P(Stats), vp]),
        "lm_reset_stats": (ctypes.c_int, [vp, vp]),
        "lm_launch_count": (i64, [vp]),
        "lm_set_option": (ctypes.c_int, [vp, i32, i64]),
        "lm_join": (ctypes.c_int, [vp, vp]),
        "lm_phase_times": (ctypes.c_int, [vp, P(flt)]),
        "lm_pair_distance_hist": (ctypes.c_int, [vp, vp, i64, flt, i32, vp, vp]),
        "lm_rasterize": (ctypes.c_int, [vp, vp, vp, i64, dbl, dbl, dbl, dbl, i32, i32, vp, vp, vp]),
        "lm_compose_frame": (ctypes.c_int, [vp, vp, vp, i32, i32, i32, ctypes.c_char_p, vp, vp]),
        "lm_record_delta_pack": (ctypes.c_int, [vp, vp, vp, vp, i64, vp, vp, vp, i64, vp, vp]),
        "lm_record_delta_unpack_host": (ctypes.c_int, [vp, vp, vp, vp, vp, i64, i64, vp, vp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    return L


EXPORTS = ["lm_version", "lm_error_string", "lm_last_cuda_error", "lm_create", "lm_destroy", "lm_set_field",
           "lm_update_field_data", "lm_set_grid", "lm_get_grid", "lm_advect_rk4", "lm_diffuse", "lm_diffuse_ids", "lm_find_pairs", "lm_interact_rps",
           "lm_pair_uniforms", "lm_resolve_rps", "lm_state_set", "lm_state_size", "lm_step", "lm_state_get",
           "lm_state_get_host", "lm_host_copies_sync", "lm_state_view", "lm_sync_stats", "lm_reset_stats", "lm_launch_count",
           "lm_phase_times", "lm_strip_alloc", "lm_set_strip", "lm_strip_buffers_get", "lm_strip_peer_export", "lm_strip_peer_connect", "lm_step_push",
           "lm_step_move", "lm_step_bin",
           "lm_step_interact_begin", "lm_step_interact_end", "lm_step_finish", "lm_set_option", "lm_join", "lm_record_next_step", "lm_record_next_step_ids", "lm_record_count",
           "lm_pair_distance_hist", "lm_rasterize", "lm_compose_frame", "lm_record_delta_pack",
           "lm_record_delta_unpack_host"]


def check(code, what):
    if code != LM_OK:
        L = lib()
        detail = L.lm_error_string(code).decode()
        if code == LM_ECUDA:
            detail += ": " + L.lm_last_cuda_error().decode()
        raise LmError(code, what, detail)
