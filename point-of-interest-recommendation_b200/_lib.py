"""ctypes binding of libpoi_b200.so -- one prototype per symbol declared in include/poi_engine.h.

There is no CPU fallback: if the shared library has not been built the import fails, and if no
B200 is present ``poi_engine_create`` fails; both loudly.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# POI_B200_LIB: an instrumented build of the same library (tools/fused_trace.py); never a different backend
LIB_PATH = os.environ.get("POI_B200_LIB") or os.path.join(_HERE, "libpoi_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libpoi_b200.so is not built (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "from the repository root. There is no CPU fallback." % LIB_PATH)



def _source_hash():
    """Same digest as __graft_entry__.source_hash(): csrc/*, include/poi_engine.h and the nvcc flags."""
    import hashlib
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]
    csrc = os.path.join(_HERE, "csrc")
    hdr = os.path.join(os.path.dirname(_HERE), "include", "poi_engine.h")
    h = hashlib.sha256(" ".join(flags).encode())
    for f in [os.path.join(csrc, f) for f in sorted(os.listdir(csrc))] + [hdr]:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


# a prebuilt binary that no longer matches its sources must not run silently (it travels to the GPU box as a file)
_stamp = LIB_PATH + ".srchash"
if not os.environ.get("POI_B200_LIB") and os.path.exists(_stamp) and os.path.isdir(os.path.join(_HERE, "csrc")):
    if open(_stamp).read().strip() != _source_hash():
        raise ImportError("libpoi_b200.so is stale: csrc/ or include/poi_engine.h changed since it was built. Rebuild with "
                          "`python -c 'import __graft_entry__ as g; g.build()'`.")

lib = ctypes.CDLL(LIB_PATH)


class PoiGruParams(Structure):
    _fields_ = [("lt", c_void_p), ("n_rows_lt", c_int64), ("d", c_int32), ("H", c_int32),
                ("ui", c_void_p), ("wh", c_void_p), ("bi", c_void_p),
                ("di", c_void_p), ("n_rows_di", c_int32), ("vs", c_void_p), ("bs", c_void_p),
                ("scal", c_void_p)]


class PoiSeqIndex(Structure):
    _fields_ = [("p", c_void_p), ("q", c_void_p), ("dp", c_void_p), ("dq", c_void_p),
                ("lens", c_void_p), ("lmax", c_int32), ("n_user", c_int32)]


class PoiGeoieParams(Structure):
    _fields_ = [("g", c_void_p), ("h", c_void_p), ("z", c_void_p), ("t", c_void_p), ("ab", c_void_p),
                ("n_rows", c_int64), ("H", c_int32)]


class PoiMgPeers(Structure):
    _fields_ = [("world", c_int32), ("rank", c_int32), ("cap", c_int64), ("n_local_rows", c_int64)] + \
               [(k, c_void_p * 16) for k in ("shard", "ob_ids", "ob_grads", "ob_cnts", "ob_perm", "ob_meta", "dense", "sums", "flags")] + \
               [("slot_tab", c_void_p)]


class PoiMfPeers(Structure):
    _fields_ = [("world", c_int32), ("rank", c_int32), ("cap", c_int64 * 2), ("n_local_rows", c_int64),
                ("shard", (c_void_p * 16) * 3), ("ob_ids", (c_void_p * 16) * 2), ("ob_perm", (c_void_p * 16) * 2),
                ("ob_meta", (c_void_p * 16) * 2), ("ob_grads", (c_void_p * 16) * 3), ("sums", c_void_p * 16),
                ("flags", c_void_p * 16), ("slot_tab", c_void_p * 2)]


_E = c_void_p
_PROTOS = {
    "poi_engine_create": (c_int, [c_int, POINTER(_E)]),
    "poi_engine_destroy": (None, [_E]),
    "poi_last_error": (c_char_p, [_E]),
    "poi_set_stream": (c_int, [_E, c_void_p]),
    "poi_sync": (c_int, [_E]),
    "poi_launch_count": (c_int, [_E, POINTER(c_int64)]),
    "poi_last_phase_ms": (c_int, [_E, POINTER(c_float)]),
    "poi_enable_phase_timing": (c_int, [_E, c_int]),
    "poi_kprof_enable": (c_int, [_E, c_int]),
    "poi_kprof_reset": (c_int, [_E]),
    "poi_kprof_get": (c_int, [_E, POINTER(c_double)]),
    "poi_set_gemm_mode": (c_int, [_E, c_int]),
    "poi_get_gemm_mode": (c_int, [_E, POINTER(c_int)]),
    "poi_set_fused_recurrence": (c_int, [_E, c_int]),
    "poi_set_fused_cluster": (c_int, [_E, c_int]),
    "poi_set_fused_sort": (c_int, [_E, c_int]),
    "poi_set_graph_mode": (c_int, [_E, c_int]),
    "poi_set_small_batch_path": (c_int, [_E, c_int]),
    "poi_graph_replays": (c_int, [_E, POINTER(c_int64)]),
    "poi_set_wgrad_mn": (c_int, [_E, c_int]),
    "poi_gather_rows": (c_int, [_E, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "poi_unique": (c_int, [_E, c_void_p, c_int64, c_int32, c_void_p, c_void_p, POINTER(c_int64)]),
    "poi_scatter_sgd": (c_int, [_E, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_float, c_float]),
    "poi_gemm_tn": (c_int, [_E, c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_int]),
    "poi_gemm_atb": (c_int, [_E, c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_int]),
    "poi_sumsq": (c_int, [_E, c_void_p, c_int64, POINTER(c_double)]),
    "poi_gru_train": (c_int, [_E, POINTER(PoiGruParams), POINTER(PoiSeqIndex), c_void_p, c_int32, c_int32,
                              c_float, c_float, POINTER(c_double)]),
    "poi_gru_train_host_rows": (c_int, [_E, POINTER(PoiGruParams), c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_int32, c_int32, c_float, c_float, POINTER(c_double)]),
    "poi_gru_predict": (c_int, [_E, POINTER(PoiGruParams), POINTER(PoiSeqIndex), c_void_p, c_int32, c_int32,
                                c_void_p, c_void_p]),
    "poi_gru_mg_dense_size": (c_int, [POINTER(PoiGruParams), POINTER(c_int64)]),
    "poi_gru_mg_prepare": (c_int, [_E, POINTER(PoiGruParams), POINTER(PoiSeqIndex), c_void_p, c_int32, c_void_p, POINTER(c_int64)]),
    "poi_gru_train_mg": (c_int, [_E, POINTER(PoiGruParams), POINTER(PoiSeqIndex), c_void_p, c_int32, c_int32, c_int32,
                                 c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "poi_gru_apply_mg": (c_int, [_E, POINTER(PoiGruParams), c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_int64,
                                 c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, POINTER(c_double)]),
    "poi_peer_alloc": (c_int, [_E, c_int64, POINTER(c_void_p), c_void_p]),
    "poi_peer_free": (c_int, [_E, c_void_p]),
    "poi_peer_open": (c_int, [_E, c_void_p, POINTER(c_void_p)]),
    "poi_peer_close": (c_int, [_E, c_void_p]),
    "poi_gather_rows_sharded": (c_int, [_E, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "poi_group_by_owner": (c_int, [_E, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "poi_pull_segments": (c_int, [_E, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "poi_sample_negatives": (c_int, [_E, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_uint64,
                                     c_uint32, c_void_p]),
    "poi_neg_intervals": (c_int, [_E, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_double, c_int32, c_void_p]),
    "poi_bpr_train_seq": (c_int, [_E, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_float, c_float, c_void_p]),
    "poi_bpr_train_batch": (c_int, [_E, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_int64, c_float, c_float, POINTER(c_double)]),
    "poi_prme_train_seq": (c_int, [_E, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_double, c_float, c_float,
                                   c_void_p]),
    "poi_prme_train_seq_k": (c_int, [_E, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_double, c_float, c_float,
                                     c_void_p]),
    "poi_prme_train_batch_k": (c_int, [_E, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_double,
                                       c_float, c_float, POINTER(c_double)]),
    "poi_gru_step_mg": (c_int, [_E, POINTER(PoiGruParams), POINTER(PoiSeqIndex), c_void_p, c_int32, c_int32, POINTER(PoiMgPeers),
                                c_int64, c_float, c_float, POINTER(c_double)]),
    "poi_gru_step_mg_host_rows": (c_int, [_E, POINTER(PoiGruParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                          POINTER(PoiMgPeers), c_int64, c_float, c_float, POINTER(c_double)]),
    "poi_geoie_step_mg": (c_int, [_E, POINTER(PoiGeoieParams), c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                  POINTER(PoiMfPeers), c_int64, c_float, c_float, POINTER(c_double)]),
    "poi_geoie_train": (c_int, [_E, POINTER(PoiGeoieParams), c_int32, c_void_p, c_void_p, c_int32, c_void_p,
                                c_void_p, c_void_p, c_int32, c_float, c_float, POINTER(c_double)]),
    "poi_geoie_train_batch_k": (c_int, [_E, POINTER(PoiGeoieParams), c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                        c_float, c_float, POINTER(c_double)]),
    "poi_score_topk_geo": (c_int, [_E, c_void_p, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_double,
                                   c_int32, c_float, c_int32, c_void_p]),
    "poi_score_topk": (c_int, [_E, c_void_p, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_float, c_int32,
                               c_void_p]),
}

EXPORTED = sorted(_PROTOS)

for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name)          # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args
