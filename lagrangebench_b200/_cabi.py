"""ctypes binding of ``include/lb200.h`` (the C-ABI shared library ``_lb200.so``).

There is no CPU fallback: if the library is missing and cannot be built, importing the
compute path raises.  torch tensors are passed as raw device pointers + the current
stream; nothing in the signatures is a torch type.
"""

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_lb200.so")

MAX_NODE_IN = 64
LATENT = 128
OVF_NEIGHBOR_LIST = 1
OVF_CELL_LIST = 2


class Grid(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_valid", C.c_int32), ("dim", C.c_int32), ("pos_f64", C.c_int32),
                ("periodic", C.c_int32), ("box", C.c_double * 3), ("r_cutoff", C.c_double), ("use_cells", C.c_int32),
                ("cells_per_side", C.c_int32 * 3), ("cell_size", C.c_float * 3), ("n_cells", C.c_int32),
                ("n_cand_cells", C.c_int32)]


class FeatureCfg(C.Structure):
    _fields_ = [("n", C.c_int32), ("dim", C.c_int32), ("t_window", C.c_int32), ("pos_f64", C.c_int32),
                ("periodic", C.c_int32), ("box", C.c_double * 3), ("r_cutoff", C.c_double),
                ("vel_mean", C.c_double * 3), ("vel_std", C.c_double * 3),
                ("magnitude_features", C.c_int32), ("bound_features", C.c_int32),
                ("bounds_lo", C.c_double * 3), ("bounds_hi", C.c_double * 3),
                ("force_mode", C.c_int32), ("force_axis", C.c_int32), ("force_threshold", C.c_double),
                ("force_lo", C.c_double * 3), ("force_hi", C.c_double * 3), ("node_stride", C.c_int32),
                ("embed_size", C.c_int32), ("num_particle_types", C.c_int32), ("embedding_dev", C.c_void_p),
                ("ptype_dev", C.c_void_p)]


class MlpOff(C.Structure):
    _fields_ = [("w0", C.c_int64), ("b0", C.c_int64), ("w1", C.c_int64), ("b1", C.c_int64),
                ("ln_scale", C.c_int64), ("ln_offset", C.c_int64), ("tc_w", C.c_int64), ("tc_vec", C.c_int64)]


class GnsCfg(C.Structure):
    _fields_ = [("n", C.c_int32), ("dim", C.c_int32), ("num_mp_steps", C.c_int32), ("node_in", C.c_int32),
                ("node_stride", C.c_int32), ("embed_size", C.c_int32), ("num_particle_types", C.c_int32),
                ("e_cap", C.c_int32), ("embedding", C.c_int64), ("enc_node", MlpOff), ("enc_edge", MlpOff),
                ("dec", MlpOff), ("proc_edge", C.POINTER(MlpOff)), ("proc_node", C.POINTER(MlpOff)),
                ("edge_impl", C.c_int32), ("n_owned", C.c_int32), ("shard", C.c_void_p),
                ("nonfinite_flag", C.c_void_p), ("node_feat_embedded", C.c_int32), ("latent", C.c_int32)]


MAX_RANKS = 16
OVF_DRIFT = 4
OVF_PEER_TIMEOUT = 8
ERR_NONFINITE = 16


class Shard(C.Structure):
    """``lb200_shard``: one rank's view of a slab decomposition (include/lb200.h)."""

    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("has_left", C.c_int32), ("has_right", C.c_int32),
                ("n_owned", C.c_int32), ("n_ghost_left", C.c_int32), ("n_ghost_right", C.c_int32),
                ("n_send_left", C.c_int32), ("n_send_right", C.c_int32),
                ("push_left", C.c_void_p), ("push_right", C.c_void_p),
                ("dst_row_left", C.c_int32), ("dst_row_right", C.c_int32), ("axis", C.c_int32), ("n_cap", C.c_int32),
                ("shift_left", C.c_double), ("shift_right", C.c_double), ("axis_length", C.c_double),
                ("ref_coord", C.c_void_p), ("drift_limit", C.c_double),
                ("heap", C.c_void_p), ("heap_left", C.c_void_p), ("heap_right", C.c_void_p),
                ("heap_all", C.c_void_p * MAX_RANKS)]


class IntegrateCfg(C.Structure):
    _fields_ = [("n", C.c_int32), ("dim", C.c_int32), ("t_window", C.c_int32), ("pos_f64", C.c_int32),
                ("periodic", C.c_int32), ("out_mode", C.c_int32), ("box", C.c_double * 3),
                ("mean", C.c_double * 3), ("std", C.c_double * 3)]


class RolloutCfg(C.Structure):
    _fields_ = [("grid", Grid), ("feat", FeatureCfg), ("gns", GnsCfg), ("integ", IntegrateCfg),
                ("cell_capacity", C.c_int32), ("e_cap", C.c_int32), ("shard", C.c_void_p)]


_VP, _I32, _I64 = C.c_void_p, C.c_int32, C.c_int64
_SIGNATURES = {
    "lb200_version": (C.c_int, []),
    "lb200_error_string": (C.c_char_p, [C.c_int]),
    "lb200_grid_init": (C.c_int, [C.POINTER(Grid), _I32, _I32, _I32, _I32, C.POINTER(C.c_double), C.c_double]),
    "lb200_nbr_scratch_bytes": (_I64, [C.POINTER(Grid)]),
    "lb200_csr_scratch_bytes": (_I64, [_I32, _I32]),
    "lb200_nbr_build": (C.c_int, [C.POINTER(Grid), _VP, _I32, _VP, _I32, _VP, _VP, _I64, _VP]),
    "lb200_nbr_csr_build": (C.c_int, [C.POINTER(Grid), _VP, _I64, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _I32, _VP, _VP, _I64,
                                      _VP]),
    "lb200_csr_build": (C.c_int, [_VP, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _I64, _VP]),
    "lb200_node_feature_width": (_I32, [C.POINTER(FeatureCfg)]),
    "lb200_features": (C.c_int, [C.POINTER(FeatureCfg), _VP, _VP, _VP, _I32, _VP, _VP, _VP]),
    "lb200_gns_scratch_bytes": (_I64, [_I32, _I32]),
    "lb200_gns_scratch_layout": (C.c_int, [_I32, _I32, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "lb200_gns_forward": (C.c_int, [C.POINTER(GnsCfg), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I64, _VP]),
    "lb200_integrate": (C.c_int, [C.POINTER(IntegrateCfg), _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "lb200_rollout_scratch_bytes": (_I64, [C.POINTER(RolloutCfg)]),
    "lb200_rollout_steps": (C.c_int, [C.POINTER(RolloutCfg), _I32, _VP, _VP, _VP, _VP, _VP, _VP, _I32, _VP, _VP, _VP, _I64,
                                      _VP]),
    "lb200_peer_heap_layout": (_I64, [_I32, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "lb200_peer_heap_create": (C.c_int, [_I64, C.POINTER(C.c_void_p), C.c_char_p]),
    "lb200_peer_heap_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "lb200_peer_heap_close": (C.c_int, [_VP]),
    "lb200_peer_heap_destroy": (C.c_int, [_VP]),
    "lb200_tc_selftest": (C.c_int, [_VP, _VP]),
    "lb200_launch_count": (_I64, []),
    "lb200_profile": (C.c_int, [_I32]),
    "lb200_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}
EXPORTED = tuple(_SIGNATURES)

_lib = None


def library_path():
    return _SO


def load():
    """Load the library, building it first when it is absent OR stale: the build stamp must equal the
    digest of the current sources (``build.build`` recompiles only then).  Raises when the sources
    changed and nvcc is not there to rebuild them -- a silently loaded old binary would void every test."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    try:
        _build.build()
    except RuntimeError as exc:
        if not os.path.exists(_SO):
            raise
        stamp = _SO + ".stamp"
        fresh = os.path.exists(stamp) and open(stamp).read().strip() == _build._digest()
        if not fresh:
            raise RuntimeError(f"{_SO} is older than its sources and cannot be rebuilt: {exc}") from exc
    lib = C.CDLL(_SO)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().lb200_error_string(int(rc))
        raise RuntimeError(f"lb200 call failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "lb200 needs contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("lagrangebench_b200 has no CPU path: a CUDA device (B200, sm_100a) is required")


def vec3(values, fill=0.0):
    out = (C.c_double * 3)(fill, fill, fill)
    for k, v in enumerate(values):
        out[k] = float(v)
    return out
