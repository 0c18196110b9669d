"""Thin Python face of the C-ABI: turns torch CUDA tensors into raw device pointers and numpy
arrays into host pointers.  PyTorch is only the device-memory container here; every kernel that
runs is in libpoi_b200.so.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_float, c_int, c_int64, c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import PoiGeoieParams, PoiGruParams, PoiMfPeers, PoiMgPeers, PoiSeqIndex, lib



class _RawDeviceBuffer:
    """Device memory the engine allocated (poi_peer_alloc) or mapped from a peer (poi_peer_open), presented to
    torch through __cuda_array_interface__ -- torch stays the container, the memory is IPC-exportable."""

    def __init__(self, engine, ptr: int, nbytes: int, owned: bool):
        self.engine, self.ptr, self.nbytes, self.owned = engine, ptr, nbytes, owned
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    def as_tensor(self, shape, dtype):
        """torch view of OWN memory (never of a peer mapping: torch would place it on the peer's device)."""
        t = torch.as_tensor(self, device=self.engine.torch_device).view(dtype)
        t = t[:int(np.prod(shape))].view(*shape)
        t._poi_raw = self           # keep the allocation alive as long as the tensor
        return t

    def view(self, shape):
        """Pointer + shape of a peer mapping, for the engine calls that read peer memory."""
        self.shape = tuple(shape)
        return self

    def data_ptr(self):
        return self.ptr

    def __del__(self):
        try:
            if self.ptr and getattr(self.engine, "_h", None):
                (lib.poi_peer_free if self.owned else lib.poi_peer_close)(self.engine._h, c_void_p(self.ptr))
        except Exception:
            pass
        self.ptr = 0


class EngineError(RuntimeError):
    pass


def _dev_f32(t: torch.Tensor, name: str) -> int:
    if t is None:
        return 0
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise TypeError("%s must be a contiguous float32 CUDA tensor" % name)
    return t.data_ptr()


def _dev_i32(t: torch.Tensor, name: str) -> int:
    if t is None:
        return 0
    if not (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
        raise TypeError("%s must be a contiguous int32 CUDA tensor" % name)
    return t.data_ptr()


def _host_i32(a, name: str) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


class Engine:
    """One engine per (process, GPU)."""

    _instances = {}

    def __init__(self, device: int | None = None):
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = int(device)
        h = ctypes.c_void_p()
        rc = lib.poi_engine_create(self.device, byref(h))
        if rc != 0:
            raise EngineError("poi_engine_create failed (%d): %s -- the B200 CUDA path is the only path; "
                              "there is no CPU fallback" % (rc, lib.poi_last_error(None).decode()))
        self._h = h
        self.torch_device = torch.device("cuda", self.device)

    @classmethod
    def get(cls, device: int | None = None) -> "Engine":
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        if device not in cls._instances:
            cls._instances[device] = Engine(device)
        return cls._instances[device]

    def close(self):
        if getattr(self, "_h", None):
            lib.poi_engine_destroy(self._h)
            self._h = None

    def _ck(self, rc: int):
        if rc != 0:
            raise EngineError(lib.poi_last_error(self._h).decode())

    # ---- plumbing ---------------------------------------------------------------------------
    def set_stream(self, stream: torch.cuda.Stream | None):
        self._ck(lib.poi_set_stream(self._h, stream.cuda_stream if stream is not None else 0))

    def sync(self):
        self._ck(lib.poi_sync(self._h))

    def launch_count(self) -> int:
        n = c_int64()
        self._ck(lib.poi_launch_count(self._h, byref(n)))
        return n.value

    def enable_phase_timing(self, on: bool):
        self._ck(lib.poi_enable_phase_timing(self._h, 1 if on else 0))

    def last_phase_ms(self):
        buf = (c_float * 8)()
        self._ck(lib.poi_last_phase_ms(self._h, buf))
        return list(buf)

    KPROF_CATS = ["other", "index", "gather", "gemm", "wgrad", "loss", "eltwise", "rows", "mf", "geoie", "eval", "reduce",
                  "recur_fwd", "recur_bwd"]

    def kprof_enable(self, on: bool):
        self._ck(lib.poi_kprof_enable(self._h, 1 if on else 0))

    def kprof_reset(self):
        self._ck(lib.poi_kprof_reset(self._h))

    def kprof_get(self):
        """{category: dict(ms, launches, flops, bytes)} accumulated since the last reset."""
        buf = (c_double * (4 * len(self.KPROF_CATS)))()
        self._ck(lib.poi_kprof_get(self._h, buf))
        return {c: dict(ms=buf[i * 4], launches=int(buf[i * 4 + 1]), flops=buf[i * 4 + 2], bytes=buf[i * 4 + 3])
                for i, c in enumerate(self.KPROF_CATS)}

    def set_gemm_mode(self, mode: int):
        self._ck(lib.poi_set_gemm_mode(self._h, int(mode)))

    def set_fused_recurrence(self, on: bool):
        self._ck(lib.poi_set_fused_recurrence(self._h, 1 if on else 0))

    def set_fused_cluster(self, cl: int):
        """CTAs per 128 users in the fused recurrence kernels: 0 auto, 1, 2 or 4."""
        self._ck(lib.poi_set_fused_cluster(self._h, int(cl)))

    def set_fused_sort(self, on: bool):
        """Index lists > 4096 keys: radix passes + segment arrays in one persistent launch (default) or one launch per phase."""
        self._ck(lib.poi_set_fused_sort(self._h, int(bool(on))))

    def set_graph_mode(self, on: bool):
        """CUDA-graph replay of train calls with B <= 8 (the reference's one-by-one mode); default on."""
        self._ck(lib.poi_set_graph_mode(self._h, 1 if on else 0))

    def set_small_batch_path(self, on: bool):
        """B <= 8: SIMT recurrence kernels with Wh resident in shared memory (default on)."""
        self._ck(lib.poi_set_small_batch_path(self._h, 1 if on else 0))

    def graph_replays(self) -> int:
        n = c_int64()
        self._ck(lib.poi_graph_replays(self._h, byref(n)))
        return n.value

    def set_wgrad_mn(self, on: bool):
        self._ck(lib.poi_set_wgrad_mn(self._h, 1 if on else 0))

    def get_gemm_mode(self) -> int:
        m = c_int()
        self._ck(lib.poi_get_gemm_mode(self._h, byref(m)))
        return m.value


    # ---- NVLink peer memory (csrc/peer.cuh) ---------------------------------------------------
    def peer_alloc(self, shape, dtype: torch.dtype):
        """Device tensor in memory other ranks of the box can map: returns (tensor, 64-byte IPC handle)."""
        n = int(np.prod(shape))
        nbytes = max(n, 1) * torch.empty((), dtype=dtype).element_size()
        ptr = c_void_p()
        handle = ctypes.create_string_buffer(64)
        self._ck(lib.poi_peer_alloc(self._h, nbytes, byref(ptr), handle))
        return _RawDeviceBuffer(self, ptr.value, nbytes, owned=True).as_tensor(shape, dtype), handle.raw

    def peer_open(self, handle: bytes, shape, dtype: torch.dtype):
        """Map a peer rank's buffer (poi_peer_alloc on that rank) into this process; returns a pointer view
        (data_ptr(), shape) for gather_rows_sharded / pull_segments."""
        n = int(np.prod(shape))
        nbytes = max(n, 1) * torch.empty((), dtype=dtype).element_size()
        ptr = c_void_p()
        self._ck(lib.poi_peer_open(self._h, ctypes.c_char_p(handle), byref(ptr)))
        return _RawDeviceBuffer(self, ptr.value, nbytes, owned=False).view(shape)

    def gather_rows_sharded(self, shards, ids: torch.Tensor, out: torch.Tensor):
        """out[i] = row ids[i] of the row-sharded table (owner = id % world), read from the owners' shards."""
        world = len(shards)
        arr = (c_void_p * world)(*[s.data_ptr() for s in shards])
        self._ck(lib.poi_gather_rows_sharded(self._h, arr, world, shards[0].shape[1], _dev_i32(ids, "ids"), ids.numel(),
                                             _dev_f32(out, "out")))
        return out

    def group_by_owner(self, ids: torch.Tensor, world: int, perm_out: torch.Tensor, counts_out: torch.Tensor):
        """perm_out = record numbers grouped by owner (ids % world), stable; counts_out (float64[world]) = group sizes."""
        self._ck(lib.poi_group_by_owner(self._h, _dev_i32(ids, "ids"), ids.numel(), world, _dev_i32(perm_out, "perm"),
                                        counts_out.data_ptr()))

    def pull_segments(self, perm, ids, grads, cnts, src_off, n, recv_ids, recv_grads, recv_cnts):
        """Copy, from every rank's outbox, the n[r] records perm[r][src_off[r] ...] (peer memory) into this rank's
        receive buffers, grouped by source rank."""
        world = len(ids)
        a_pm = (c_void_p * world)(*[t.data_ptr() for t in perm])
        a_ids = (c_void_p * world)(*[t.data_ptr() for t in ids])
        a_gr = (c_void_p * world)(*[t.data_ptr() for t in grads])
        a_cn = (c_void_p * world)(*[t.data_ptr() for t in cnts])
        a_so = (c_int64 * world)(*[int(x) for x in src_off])
        a_n = (c_int64 * world)(*[int(x) for x in n])
        self._ck(lib.poi_pull_segments(self._h, world, grads[0].shape[1], a_pm, a_ids, a_gr, a_cn, a_so, a_n,
                                       _dev_i32(recv_ids, "recv_ids"), _dev_f32(recv_grads, "recv_grads"),
                                       _dev_f32(recv_cnts, "recv_cnts")))


    # ---- per-epoch negative sampling on the device (csrc/sampling.cuh) ---------------
    def sample_negatives(self, rows: torch.Tensor, sorted_a: torch.Tensor, n_item: int, seed: int, epoch: int,
                         sorted_b: torch.Tensor | None = None, out: torch.Tensor | None = None):
        """out[u, t] = uniform draw from [0, n_item) not in the user's sorted rows, for positions before the first pad."""
        if out is None:
            out = torch.empty_like(rows)
        self._ck(lib.poi_sample_negatives(self._h, _dev_i32(rows, "rows"), rows.shape[1], _dev_i32(sorted_a, "sorted_a"),
                                          sorted_a.shape[1], _dev_i32(sorted_b, "sorted_b") if sorted_b is not None else None,
                                          sorted_b.shape[1] if sorted_b is not None else 0, rows.shape[0], int(n_item),
                                          int(seed) & 0xFFFFFFFFFFFFFFFF, int(epoch) & 0xFFFFFFFF, _dev_i32(out, "out")))
        return out

    def neg_intervals(self, p: torch.Tensor, q: torch.Tensor, lens: torch.Tensor, coords: torch.Tensor, dd: float,
                      dist_num: int, out: torch.Tensor | None = None):
        """Distance-interval ids between q[u, t] and p[u, t-1] (cal_dis in fp64); dist_num at t = 0 and on the padding."""
        if coords.dtype != torch.float64 or not coords.is_contiguous():
            raise EngineError("coords must be a contiguous float64 [n_item x 2] tensor")
        if out is None:
            out = torch.empty_like(p)
        self._ck(lib.poi_neg_intervals(self._h, _dev_i32(p, "p"), _dev_i32(q, "q"), _dev_i32(lens, "lens"), p.shape[0], p.shape[1],
                                       coords.data_ptr(), float(dd), int(dist_num), _dev_i32(out, "out")))
        return out

    # ---- first-slice kernels ----------------------------------------------------------------
    def gather_rows(self, table: torch.Tensor, idx: torch.Tensor, out: torch.Tensor | None = None):
        n = idx.numel()
        if out is None:
            out = torch.empty((n, table.shape[1]), dtype=torch.float32, device=table.device)
        self._ck(lib.poi_gather_rows(self._h, _dev_f32(table, "table"), table.shape[0], table.shape[1],
                                     _dev_i32(idx, "idx"), n, _dev_f32(out, "out")))
        return out

    def unique(self, idx: torch.Tensor, key_bound: int):
        n = idx.numel()
        uq = torch.empty(max(n, 1), dtype=torch.int32, device=idx.device)
        cnt = torch.empty(max(n, 1), dtype=torch.int32, device=idx.device)
        nu = c_int64()
        self._ck(lib.poi_unique(self._h, _dev_i32(idx, "idx"), n, int(key_bound), uq.data_ptr(), cnt.data_ptr(), byref(nu)))
        return uq[: nu.value], cnt[: nu.value]

    def scatter_sgd(self, table: torch.Tensor, idx: torch.Tensor, grad: torch.Tensor | None, alpha: float, lam: float):
        self._ck(lib.poi_scatter_sgd(self._h, _dev_f32(table, "table"), table.shape[0], table.shape[1],
                                     _dev_i32(idx, "idx"), idx.numel(), _dev_f32(grad, "grad"), alpha, lam))

    def gemm_tn(self, A: torch.Tensor, W: torch.Tensor, bias: torch.Tensor | None = None, mode: int = 0):
        """C = A @ W.T (+ bias); A [M, K], W [N, K] float32 CUDA, row-major."""
        M, K = A.shape
        N = W.shape[0]
        ldc = (N + 3) // 4 * 4
        C = torch.empty((M, ldc), dtype=torch.float32, device=A.device)
        self._ck(lib.poi_gemm_tn(self._h, _dev_f32(A, "A"), K, _dev_f32(W, "W"), K, M, N, K, _dev_f32(bias, "bias"),
                                 C.data_ptr(), ldc, int(mode)))
        return C[:, :N]

    def gemm_atb(self, A: torch.Tensor, B: torch.Tensor, mode: int = 0):
        """C = A.T @ B; A [M, N1], B [M, N2] float32 CUDA row-major (N1, N2 multiples of 4)."""
        C = torch.empty((A.shape[1], B.shape[1]), dtype=torch.float32, device=A.device)
        self._ck(lib.poi_gemm_atb(self._h, _dev_f32(A, "A"), A.shape[1], _dev_f32(B, "B"), B.shape[1], A.shape[0],
                                  A.shape[1], B.shape[1], C.data_ptr(), int(mode)))
        return C

    def sumsq(self, x: torch.Tensor) -> float:
        out = c_double()
        self._ck(lib.poi_sumsq(self._h, _dev_f32(x, "x"), x.numel(), byref(out)))
        return out.value

    # ---- GRU family -------------------------------------------------------------------------
    @staticmethod
    def gru_params(lt, ui, wh, bi, di=None, vs=None, bs=None, scal=None) -> PoiGruParams:
        p = PoiGruParams()
        p.lt = _dev_f32(lt, "lt"); p.n_rows_lt = lt.shape[0]; p.d = lt.shape[1]; p.H = wh.shape[1]
        p.ui = _dev_f32(ui, "ui"); p.wh = _dev_f32(wh, "wh"); p.bi = _dev_f32(bi, "bi")
        if di is not None:
            p.di = _dev_f32(di, "di"); p.n_rows_di = di.shape[0]
            p.vs = _dev_f32(vs, "vs"); p.bs = _dev_f32(bs, "bs"); p.scal = _dev_f32(scal, "scal")
        else:
            p.di = 0; p.n_rows_di = 0; p.vs = 0; p.bs = 0; p.scal = 0
        return p

    @staticmethod
    def seq_index(p, q, lens, dp=None, dq=None) -> PoiSeqIndex:
        ix = PoiSeqIndex()
        ix.p = _dev_i32(p, "p"); ix.q = _dev_i32(q, "q"); ix.lens = _dev_i32(lens, "lens")
        ix.dp = _dev_i32(dp, "dp"); ix.dq = _dev_i32(dq, "dq")
        ix.n_user, ix.lmax = p.shape
        return ix

    def gru_train(self, params: PoiGruParams, index: PoiSeqIndex, uidx, max_len: int, alpha: float, lam: float):
        uidx = _host_i32(uidx, "uidx").reshape(-1)
        out = (c_double * 5)()
        self._ck(lib.poi_gru_train(self._h, byref(params), byref(index), uidx.ctypes.data, uidx.size, int(max_len),
                                   alpha, lam, out))
        return list(out)

    def gru_train_host_rows(self, params: PoiGruParams, p, q, lens, alpha: float, lam: float, dp=None, dq=None):
        """p, q, dp, dq: host int32 [B, lmax] (numpy, or pinned CPU torch tensors); lens int32 [B]."""
        def hp(a):
            if a is None:
                return 0, None
            if isinstance(a, torch.Tensor):
                assert a.dtype == torch.int32 and a.is_contiguous() and not a.is_cuda
                return a.data_ptr(), a
            a = np.ascontiguousarray(a, dtype=np.int32)
            return a.ctypes.data, a
        pp, k0 = hp(p); qq, k1 = hp(q); dpp, k2 = hp(dp); dqq, k3 = hp(dq); ll, k4 = hp(lens)
        B, lmax = p.shape
        out = (c_double * 5)()
        self._ck(lib.poi_gru_train_host_rows(self._h, byref(params), pp, qq, dpp, dqq, ll, int(B), int(lmax),
                                             alpha, lam, out))
        return list(out)

    def gru_predict(self, params: PoiGruParams, index: PoiSeqIndex, uidx, max_len: int):
        uidx = _host_i32(uidx, "uidx").reshape(-1)
        B = uidx.size
        dev = self.torch_device
        hts = torch.empty((B, params.H), dtype=torch.float32, device=dev)
        sts = torch.empty((B, params.n_rows_di), dtype=torch.float32, device=dev) if params.di else None
        self._ck(lib.poi_gru_predict(self._h, byref(params), byref(index), uidx.ctypes.data, B, int(max_len),
                                     hts.data_ptr(), sts.data_ptr() if sts is not None else 0))
        return hts, sts

    # ---- multi-GPU step (two phases around the collectives) ---------------------------------
    @staticmethod
    def gru_mg_dense_size(params: PoiGruParams) -> int:
        n = c_int64()
        lib.poi_gru_mg_dense_size(byref(params), byref(n))
        return n.value

    def gru_mg_prepare(self, params, index, uidx, lmax):
        """Sorted unique row ids of the batch (int32 CUDA tensor); the next gru_train_mg reuses the sort."""
        uidx = _host_i32(uidx, "uidx").reshape(-1)
        buf = torch.empty(2 * uidx.size * lmax, dtype=torch.int32, device=self.torch_device)
        n = c_int64()
        self._ck(lib.poi_gru_mg_prepare(self._h, byref(params), byref(index), uidx.ctypes.data, uidx.size, buf.data_ptr(), byref(n)))
        return buf[: n.value]

    def gru_train_mg(self, params, index, uidx, max_len, global_batch, rows, dense_grads, row_grads, row_cnt, loss_sums):
        uidx = _host_i32(uidx, "uidx").reshape(-1)
        self._ck(lib.poi_gru_train_mg(self._h, byref(params), byref(index), uidx.ctypes.data, uidx.size, int(max_len),
                                      int(global_batch), _dev_f32(rows, "rows"), rows.shape[0],
                                      _dev_f32(dense_grads, "dense_grads"), _dev_f32(row_grads, "row_grads"),
                                      _dev_f32(row_cnt, "row_cnt"), loss_sums.data_ptr()))

    def gru_apply_mg(self, params, dense_grads, loss_sums, global_batch, n_nonempty_global, lt_local, recv_ids,
                     recv_grads, recv_cnts, alpha, lam):
        out = (c_double * 5)()
        n_recv = recv_ids.numel()
        self._ck(lib.poi_gru_apply_mg(self._h, byref(params), _dev_f32(dense_grads, "dense_grads"), loss_sums.data_ptr(),
                                      int(global_batch), int(n_nonempty_global), _dev_f32(lt_local, "lt_local"),
                                      lt_local.shape[0], _dev_i32(recv_ids, "recv_ids") if n_recv else 0,
                                      _dev_f32(recv_grads, "recv_grads") if n_recv else 0,
                                      _dev_f32(recv_cnts, "recv_cnts") if n_recv else 0, n_recv, alpha, lam, out))
        return list(out)

    def gru_step_mg(self, params, index, uidx, max_len, peers, step, alpha, lam):
        """The whole multi-GPU step in one call (csrc/mg_step.cuh); `peers` is a filled PoiMgPeers."""
        uidx = _host_i32(uidx, "uidx").reshape(-1)
        out = (c_double * 5)()
        self._ck(lib.poi_gru_step_mg(self._h, byref(params), byref(index), uidx.ctypes.data, uidx.size, int(max_len),
                                     byref(peers), int(step), alpha, lam, out))
        return list(out)

    def gru_step_mg_host_rows(self, params, p, q, dp, dq, lens, peers, step, alpha, lam):
        """Same step, index rows from (pinned) host tensors / arrays [B x lmax]: copied inside the call."""
        f = lambda x: None if x is None else np.ascontiguousarray(x.numpy() if isinstance(x, torch.Tensor) else x, dtype=np.int32)
        p, q, dp, dq, lens = f(p), f(q), f(dp), f(dq), f(lens)
        out = (c_double * 5)()
        self._ck(lib.poi_gru_step_mg_host_rows(self._h, byref(params), p.ctypes.data, q.ctypes.data,
                                               dp.ctypes.data if dp is not None else None, dq.ctypes.data if dq is not None else None,
                                               lens.ctypes.data, p.shape[0], p.shape[1], byref(peers), int(step), alpha, lam, out))
        return list(out)

    def geoie_step_mg(self, ab, n_rows_global, H, P, Q, coords, peers, step, alpha, lam) -> float:
        """One row-sharded multi-GPU GeoIE mini-batch step (csrc/mf_mg.cuh); P, Q: global row ids of this rank's users."""
        prm = PoiGeoieParams()
        prm.g = 0; prm.h = 0; prm.z = 0; prm.t = 0
        prm.ab = ab.data_ptr(); prm.n_rows = int(n_rows_global); prm.H = int(H)
        on_dev = isinstance(P, torch.Tensor) and P.is_cuda
        if on_dev:
            Bu, L = int(P.shape[0]), int(P.shape[1]); K = int(Q.shape[2])
            pp, qp = _dev_i32(P, "P"), _dev_i32(Q, "Q")
            keep = None
        else:
            f = lambda x: np.ascontiguousarray(x.numpy() if isinstance(x, torch.Tensor) else x, dtype=np.int32)
            keep = (f(P), f(Q))
            Bu, L = keep[0].shape; K = keep[1].shape[2]
            pp, qp = keep[0].ctypes.data, keep[1].ctypes.data
        out = c_double()
        self._ck(lib.poi_geoie_step_mg(self._h, byref(prm), pp, qp, _dev_f32(coords, "coords"), Bu, L, K, 0 if on_dev else 1,
                                       byref(peers), int(step), alpha, lam, byref(out)))
        return float(out.value)

    # ---- BPR / PRME -------------------------------------------------------------------------
    def bpr_train_seq(self, ux, lt, u, p, q, alpha, lam) -> np.ndarray:
        u = _host_i32(u, "u").reshape(-1); p = _host_i32(p, "p").reshape(-1); q = _host_i32(q, "q").reshape(-1)
        loss = np.empty(u.size, dtype=np.float64)
        self._ck(lib.poi_bpr_train_seq(self._h, _dev_f32(ux, "ux"), _dev_f32(lt, "lt"), lt.shape[1],
                                       u.ctypes.data, p.ctypes.data, q.ctypes.data, u.size, alpha, lam,
                                       loss.ctypes.data))
        return loss

    def bpr_train_batch(self, ux, lt, p, q, mask, u, alpha, lam) -> float:
        p = _host_i32(p, "p").reshape(-1); q = _host_i32(q, "q").reshape(-1)
        mask = _host_i32(mask, "mask").reshape(-1); u = _host_i32(u, "u").reshape(-1)
        out = c_double()
        self._ck(lib.poi_bpr_train_batch(self._h, _dev_f32(ux, "ux"), ux.shape[0], _dev_f32(lt, "lt"), lt.shape[0],
                                         lt.shape[1], p.ctypes.data, q.ctypes.data, mask.ctypes.data, u.ctypes.data,
                                         p.size, alpha, lam, byref(out)))
        return out.value

    def prme_train_seq(self, du, dp, ds, u, p, q, prev, dist, gap, thd, cw, alpha, lam) -> np.ndarray:
        u = _host_i32(u, "u").reshape(-1); p = _host_i32(p, "p").reshape(-1); q = _host_i32(q, "q").reshape(-1)
        prev = _host_i32(prev, "prev").reshape(-1); gap = _host_i32(gap, "gap").reshape(-1)
        dist = np.ascontiguousarray(dist, dtype=np.float64).reshape(-1)
        loss = np.empty(u.size, dtype=np.float64)
        self._ck(lib.poi_prme_train_seq(self._h, _dev_f32(du, "du"), _dev_f32(dp, "dp"), _dev_f32(ds, "ds"),
                                        dp.shape[1], u.ctypes.data, p.ctypes.data, q.ctypes.data, prev.ctypes.data,
                                        dist.ctypes.data, gap.ctypes.data, u.size, int(thd), float(cw), alpha, lam,
                                        loss.ctypes.data))
        return loss

    def prme_train_seq_k(self, du, dp, ds, u, p, Q, prev, dist, gap, thd, cw, alpha, lam) -> np.ndarray:
        """n sequential K-negative steps (parity mode); Q is [n, K]."""
        u = _host_i32(u, "u").reshape(-1); p = _host_i32(p, "p").reshape(-1)
        Q = _host_i32(Q, "Q").reshape(u.size, -1)
        prev = _host_i32(prev, "prev").reshape(-1); gap = _host_i32(gap, "gap").reshape(-1)
        dist = np.ascontiguousarray(dist, dtype=np.float64).reshape(-1)
        loss = np.empty(u.size, dtype=np.float64)
        self._ck(lib.poi_prme_train_seq_k(self._h, _dev_f32(du, "du"), _dev_f32(dp, "dp"), _dev_f32(ds, "ds"),
                                          dp.shape[1], u.ctypes.data, p.ctypes.data, Q.ctypes.data, prev.ctypes.data,
                                          dist.ctypes.data, gap.ctypes.data, u.size, Q.shape[1], int(thd), float(cw), alpha, lam,
                                          loss.ctypes.data))
        return loss

    def prme_train_batch_k(self, du, dp, ds, u, p, Q, prev, dist, gap, thd, cw, alpha, lam) -> float:
        """One K-negative mini-batch step (throughput mode).  The six index arrays are either all int32 / float32 CUDA
        tensors (device resident) or all host arrays / pinned CPU tensors (copied inside the call: the e2e path)."""
        on_dev = isinstance(u, torch.Tensor) and u.is_cuda
        if on_dev:
            n, K = int(u.numel()), int(Q.numel() // max(u.numel(), 1))
            ptr = [_dev_i32(u, "u"), _dev_i32(p, "p"), _dev_i32(Q, "Q"), _dev_i32(prev, "prev"), _dev_f32(dist, "dist"),
                   _dev_i32(gap, "gap")]
            keep = None
        else:
            f = lambda x, dt: np.ascontiguousarray(x.numpy() if isinstance(x, torch.Tensor) else x, dtype=dt)
            keep = [f(u, np.int32).reshape(-1), f(p, np.int32).reshape(-1), f(Q, np.int32), f(prev, np.int32).reshape(-1),
                    f(dist, np.float32).reshape(-1), f(gap, np.int32).reshape(-1)]
            n = keep[0].size; K = keep[2].size // max(n, 1)
            ptr = [a.ctypes.data for a in keep]
        out = c_double()
        self._ck(lib.poi_prme_train_batch_k(self._h, _dev_f32(du, "du"), du.shape[0], _dev_f32(dp, "dp"), _dev_f32(ds, "ds"),
                                            dp.shape[0], dp.shape[1], ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], ptr[5], n, K,
                                            0 if on_dev else 1, int(thd), float(cw), alpha, lam, byref(out)))
        return float(out.value)

    # ---- GeoIE ------------------------------------------------------------------------------
    def geoie_train_batch_k(self, g, h, z, t, ab, P, Q, coords, alpha, lam) -> float:
        """One K-negative GeoIE mini-batch step.  P [Bu, L], Q [Bu, L, K]: int32 CUDA tensors (resident) or host arrays /
        pinned CPU tensors (copied inside the call); coords float32 CUDA [n_rows, 2]."""
        prm = PoiGeoieParams()
        prm.g = _dev_f32(g, "g"); prm.h = _dev_f32(h, "h"); prm.z = _dev_f32(z, "z"); prm.t = _dev_f32(t, "t")
        prm.ab = ab.data_ptr(); prm.n_rows = g.shape[0]; prm.H = g.shape[1]
        on_dev = isinstance(P, torch.Tensor) and P.is_cuda
        if on_dev:
            Bu, L = int(P.shape[0]), int(P.shape[1]); K = int(Q.shape[2])
            pp, qp = _dev_i32(P, "P"), _dev_i32(Q, "Q")
            keep = None
        else:
            f = lambda x: np.ascontiguousarray(x.numpy() if isinstance(x, torch.Tensor) else x, dtype=np.int32)
            keep = (f(P), f(Q))
            Bu, L = keep[0].shape; K = keep[1].shape[2]
            pp, qp = keep[0].ctypes.data, keep[1].ctypes.data
        out = c_double()
        self._ck(lib.poi_geoie_train_batch_k(self._h, byref(prm), pp, qp, _dev_f32(coords, "coords"), Bu, L, K, 0 if on_dev else 1,
                                             alpha, lam, byref(out)))
        return float(out.value)

    def geoie_train(self, g, h, z, t, ab, uidx, p_row, q_row, dist_pos, dist_neg, msk, alpha, lam) -> float:
        prm = PoiGeoieParams()
        prm.g = _dev_f32(g, "g"); prm.h = _dev_f32(h, "h"); prm.z = _dev_f32(z, "z"); prm.t = _dev_f32(t, "t")
        if not (ab.is_cuda and ab.dtype == torch.float64 and ab.numel() == 2):
            raise TypeError("ab must be a float64 CUDA tensor of 2 elements")
        prm.ab = ab.data_ptr(); prm.n_rows = g.shape[0]; prm.H = g.shape[1]
        p_row = _host_i32(p_row, "p_row").reshape(-1); q_row = _host_i32(q_row, "q_row").reshape(-1)
        msk = _host_i32(msk, "msk")
        n = msk.shape[0]
        dpos = np.ascontiguousarray(dist_pos, dtype=np.float32); dneg = np.ascontiguousarray(dist_neg, dtype=np.float32)
        out = c_double()
        self._ck(lib.poi_geoie_train(self._h, byref(prm), int(uidx), p_row.ctypes.data, q_row.ctypes.data,
                                     p_row.size, dpos.ctypes.data, dneg.ctypes.data, msk.ctypes.data, int(n),
                                     alpha, lam, byref(out)))
        return out.value

    # ---- evaluation -------------------------------------------------------------------------
    def score_topk_geo(self, users, items, top_k, sts, user_coords, item_coords, dd, dist_num, wd):
        """Distance2Pre scores + top-K with the interval probabilities looked up on the fly (no U x I matrices)."""
        B = users.shape[0]
        out = torch.empty((B, top_k), dtype=torch.int32, device=users.device)
        if user_coords.dtype != torch.float64 or item_coords.dtype != torch.float64:
            raise EngineError("coordinates must be float64")
        self._ck(lib.poi_score_topk_geo(self._h, _dev_f32(users, "users"), B, _dev_f32(items, "items"), items.shape[0], items.shape[1],
                                        _dev_f32(sts, "sts"), sts.shape[1], user_coords.contiguous().data_ptr(),
                                        item_coords.contiguous().data_ptr(), float(dd), int(dist_num), float(wd), int(top_k), out.data_ptr()))
        return out

    def score_topk(self, users, items, top_k, prob=None, wd=0.0):
        B = users.shape[0]
        out = torch.empty((B, top_k), dtype=torch.int32, device=users.device)
        self._ck(lib.poi_score_topk(self._h, _dev_f32(users, "users"), B, _dev_f32(items, "items"), items.shape[0],
                                    items.shape[1], _dev_f32(prob, "prob"), float(wd), int(top_k), out.data_ptr()))
        return out
