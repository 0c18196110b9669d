"""Multi-GPU mini-batch training: one process per GPU, `torch.distributed`
(NCCL over NVLink / NVSwitch) as the plumbing.

Partitioning
  * users (and their index-matrix rows) are split over ranks -- each rank trains its own B users per
    step, the step's loss is normalised by the global batch;
  * the item table `lt` is ROW-SHARDED, owner(row) = row % world, local index = row // world
    (interleaved, so Zipf-hot rows spread evenly);
  * dense weights (ui, wh, bi, vs, bs, wd, loss_weight) and the small interval table `di` are
    replicated and kept identical by all-reducing their gradients.

Per step (the reference has no counterpart -- it is single process):
  1. sorted-unique row ids of the local batch (engine, bit-exact integer work);
  2. all-to-all: ids -> owners, owners gather their rows (engine), all-to-all rows back;
  3. forward + backward on the fetched rows (engine, `poi_gru_train_mg`): emits dense gradients and one
     duplicate-summed gradient row + occurrence count per unique id;
  4. all-reduce(dense gradients, loss sums); all-to-all (gradient rows, counts) -> owners;
  5. every rank applies the dense SGD step; each owner applies the sparse SGD step to its shard
     (`poi_gru_apply_mg`).
G ranks x batch B is the same update as 1 rank x batch G*B up to floating-point summation order.

`RowExchange` is device-agnostic (CPU tensors + gloo work too) so the routing logic is testable
without GPUs.

Peer-memory path (default on GPUs, `peer=True`): the whole step is ONE engine call (`poi_gru_step_mg`,
csrc/mg_step.cuh) with no collective library on its path and no host synchronisation before the final read-back.
Every rank's shard, gradient outbox, dense-gradient buffer and flag array live in IPC-exported device memory that all
ranks of the box map once.  Rows are gathered by a kernel reading the OWNERS' shards through NVLink; ranks tell each
other "my outbox is written" / "I have applied the step" by storing step counters into each other's flag arrays
(waiting kernels time out into an error instead of hanging); the dense gradients are summed in rank order by every
rank itself reading every peer's buffer (bit-identical weights everywhere); each owner finds the records addressed to
it through a direct-address table and sums them, sources in rank order, straight out of the peers' outboxes.  NCCL is
used at construction only (exchange of the IPC handles, one barrier).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class RowExchange:
    """Routing of unique row ids to their owners and back.  Built once per step from the sorted unique
    ids this rank needs; `fetch` pulls the rows, `push` sends per-row payloads (gradients, counts) to
    the owners.  Deterministic: payloads arrive grouped by source rank, in ascending id order."""

    def __init__(self, uniq_ids: torch.Tensor, world: int, group=None):
        self.world, self.group = world, group
        self.n = uniq_ids.numel()
        dev = uniq_ids.device
        ids = uniq_ids.to(torch.int64)
        owner = ids % world
        # stable sort by owner keeps ascending id order inside each destination bucket
        self.order = torch.sort(owner, stable=True).indices
        self.send_ids = ids[self.order].contiguous()
        self.send_counts = torch.bincount(owner, minlength=world).to(torch.int64)
        if world > 1:
            rc = torch.empty(world, dtype=torch.int64, device=dev)
            dist.all_to_all_single(rc, self.send_counts, group=group)
            self.recv_counts = rc
        else:
            self.recv_counts = self.send_counts.clone()
        both = torch.stack((self.send_counts, self.recv_counts)).cpu()                     # one host sync
        self.send_split, self.recv_split = both[0].tolist(), both[1].tolist()
        self.n_recv = int(sum(self.recv_split))
        self.recv_ids = self._a2a(self.send_ids, self.send_split, self.recv_split)       # global ids I own
        self.recv_local = (self.recv_ids // world).to(torch.int32).contiguous()

    def _a2a(self, send: torch.Tensor, send_split, recv_split) -> torch.Tensor:
        out = torch.empty((int(sum(recv_split)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        if self.world > 1:
            dist.all_to_all_single(out, send.contiguous(), output_split_sizes=recv_split, input_split_sizes=send_split,
                                   group=self.group)
        else:
            out.copy_(send)
        return out

    def fetch(self, gather_local) -> torch.Tensor:
        """gather_local(local_ids int32 [m]) -> rows [m, d] of this rank's shard.  Returns the rows of
        `uniq_ids`, in that order."""
        mine = gather_local(self.recv_local)
        back = self._a2a(mine, self.recv_split, self.send_split)
        rows = torch.empty_like(back)
        rows[self.order] = back
        return rows

    def push(self, payload: torch.Tensor) -> torch.Tensor:
        """payload [n, ...] aligned with `uniq_ids` -> rows received by the owner, aligned with
        `recv_local` / `recv_ids`."""
        return self._a2a(payload[self.order].contiguous(), self.send_split, self.recv_split)


def pull_plan(count_matrix, me: int):
    """Peer-memory exchange, owner side.  count_matrix[p, o] = number of records rank p addresses to owner o; rank p's
    permutation list groups its record numbers by owner in owner order 0..W-1.  Returns, for owner `me`, the first
    entry of its group in every rank's list and the group sizes -- the arguments of `poi_pull_segments`."""
    cm = np.asarray(count_matrix, dtype=np.int64)
    world = cm.shape[0]
    return [int(cm[p, :me].sum()) for p in range(world)], [int(cm[p, me]) for p in range(world)]


def shard_rows(table: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Rows owned by `rank`: table[rank::world] (local index = global // world)."""
    return np.ascontiguousarray(table[rank::world])


def unshard_rows(shards, n_rows: int) -> np.ndarray:
    out = np.empty((n_rows,) + shards[0].shape[1:], dtype=shards[0].dtype)
    for r, s in enumerate(shards):
        out[r::len(shards)] = s
    return out


class _PhaseTrace:
    """POI_MG_TRACE=1: CUDA-event time stamps between the phases of a step (stream time, not host time)."""

    def __init__(self):
        self.acc, self.n, self.cur = {}, 0, []

    def mark(self, name):
        ev = torch.cuda.Event(enable_timing=True); ev.record(); self.cur.append((name, ev))

    def end_step(self):
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(self.cur[:-1], self.cur[1:]):
            self.acc[n1] = self.acc.get(n1, 0.0) + e0.elapsed_time(e1)
        self.cur = []; self.n += 1

    def report(self):
        return {k: round(v / max(self.n, 1), 4) for k, v in self.acc.items()}


class ShardedSpatialGru:
    """Mini-batch Distance2Pre (or plain GRU when `dist_masks` is None) over `world` GPUs.

    Index matrices hold THIS rank's users only ([n_local_user x lmax]); `init['lt']` may be the full
    table (it is sharded here) or already this rank's shard (pass lt_is_shard=True)."""

    def __init__(self, train, dist_masks, alpha_lambda, n_item, n_dist, n_in, n_hidden, init, rank=None, world=None,
                 device=None, lt_is_shard=False, group=None, peer=True, max_batch=None, uniform_batch=True):
        from .engine import Engine
        from .public.GRU import _lens_from_masks
        from .shared import Shared
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.group = group
        self.engine = Engine.get(device)
        dev = self.engine.torch_device
        self.n_item, self.n_rows = n_item, n_item + 1
        self.head = dist_masks is not None
        P, M, Q = train
        self.P, self.Q = Shared(P, "int32", dev), Shared(Q, "int32", dev)
        self._lens_host = _lens_from_masks(M, "tra_masks")
        self._lens = torch.from_numpy(self._lens_host).to(dev)
        if self.head:
            self.DP, self.DQ = Shared(dist_masks[0], "int32", dev), Shared(dist_masks[1], "int32", dev)
        self._alpha, self._lambda = float(alpha_lambda[0]), float(alpha_lambda[1])
        lt = init["lt"]
        if not lt_is_shard:
            lt = lt[self.rank::self.world] if isinstance(lt, torch.Tensor) else shard_rows(np.asarray(lt), self.rank, self.world)
        self.uniform_batch = bool(uniform_batch)      # every rank passes the same number of users per step
        self.peer = bool(peer) and self.world > 1 and dev.type == "cuda"
        if self.peer:
            # the shard lives in IPC-exportable memory so that the other ranks can read it through NVLink
            shard_t, self._h_lt = self.engine.peer_alloc(tuple(lt.shape), torch.float32)
            shard_t.copy_(lt if isinstance(lt, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(lt, dtype=np.float32)))
            lt = shard_t
        self.lt_local = Shared(lt, "float32", dev)
        f = lambda k: Shared(init[k], "float32", dev)
        self.ui, self.wh, self.bi = f("ui"), f("wh"), f("bi")
        if self.head:
            self.di, self.vs, self.bs = f("di"), f("vs"), f("bs")
            lw = np.asarray(init["loss_weight"], dtype=np.float32)
            self._scal = Shared(np.array([float(np.asarray(init["wd"])), lw[0], lw[1]], dtype=np.float32), "float32", dev)
        self.d = self.lt_local.t.shape[1]
        self._dense = torch.zeros(Engine.gru_mg_dense_size(self._params()), dtype=torch.float32, device=dev)
        self._sums = torch.zeros(3, dtype=torch.float64, device=dev)
        self.last_exchange_rows = 0
        import os
        self.trace = _PhaseTrace() if os.environ.get("POI_MG_TRACE") and dev.type == "cuda" else None
        if self.peer:
            self._setup_peer(int(max_batch) if max_batch else min(int(self.P.t.shape[0]), 4096))

    def _setup_peer(self, max_batch):
        """Allocate this rank's peer-visible buffers, exchange the IPC handles once, map every peer's buffers and fill the
        pointer table `poi_gru_step_mg` takes."""
        from ._lib import PoiMgPeers
        from .engine import Engine
        eng, W, d, dev = self.engine, self.world, self.d, self.engine.torch_device
        self._cap = 2 * max_batch * int(self.P.t.shape[1])       # unique row ids of a step: at most 2 * B * lmax
        n_dense = Engine.gru_mg_dense_size(self._params())
        spec = dict(ob_ids=((self._cap,), torch.int32), ob_grads=((self._cap, d), torch.float32), ob_cnts=((self._cap,), torch.float32),
                    ob_perm=((self._cap,), torch.int32), ob_meta=((W + 2,), torch.int32), dense=((n_dense,), torch.float32),
                    sums=((4,), torch.float64), flags=((2 * W,), torch.int32))
        own, handles = {}, {}
        for k, (shape, dt) in spec.items():
            own[k], handles[k] = eng.peer_alloc(shape, dt)
            own[k].zero_()
        own["shard"], handles["shard"] = self.lt_local.t, self._h_lt
        self._own = own
        self._slot_tab = torch.full((int(self.lt_local.t.shape[0]) * W,), -1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        mine = dict(h=handles, lt_shape=tuple(self.lt_local.t.shape), cap=self._cap)
        every = [None] * W
        dist.all_gather_object(every, mine, group=self.group)
        if any(x["cap"] != self._cap for x in every):
            raise ValueError("all ranks must use the same max_batch / lmax")
        pt = PoiMgPeers()
        pt.world, pt.rank, pt.cap, pt.n_local_rows = W, self.rank, self._cap, int(self.lt_local.t.shape[0])
        self._mapped = []
        for r, x in enumerate(every):
            for k in list(spec) + ["shard"]:
                if r == self.rank:
                    ptr = own[k].data_ptr()
                else:
                    shape, dt = (x["lt_shape"], torch.float32) if k == "shard" else spec[k]
                    m = eng.peer_open(x["h"][k], shape, dt)
                    self._mapped.append(m)
                    ptr = m.data_ptr()
                getattr(pt, k)[r] = ptr
        pt.slot_tab = self._slot_tab.data_ptr()
        self._peers = pt
        self._step_no = 0
        dist.barrier(group=self.group)

    def _params(self):
        from .engine import Engine
        p = Engine.gru_params(self.lt_local.t, self.ui.t, self.wh.t, self.bi.t,
                              self.di.t if self.head else None, self.vs.t if self.head else None,
                              self.bs.t if self.head else None, self._scal.t if self.head else None)
        p.n_rows_lt = self.n_rows          # key bound of the batch's row ids is the GLOBAL row count
        return p

    def _index(self):
        from .engine import Engine
        return Engine.seq_index(self.P.t, self.Q.t, self._lens, self.DP.t if self.head else None,
                                self.DQ.t if self.head else None)

    def train_host_rows(self, p, q, dp, dq, lens):
        """End-to-end variant: the step's index rows come from (pinned) host tensors [B x lmax]; they are
        copied to the device inside the call, then the same step runs on them."""
        from .engine import Engine
        dev = self.engine.torch_device
        B = p.shape[0]
        if self.peer:
            if 2 * B * int(p.shape[1]) > self._cap:
                raise ValueError("batch of %d users exceeds the outbox capacity: construct with a larger max_batch" % B)
            self._step_no += 1
            out = self.engine.gru_step_mg_host_rows(self._params(), p, q, dp if self.head else None, dq if self.head else None, lens,
                                                    self._peers, self._step_no, self._alpha, self._lambda)
            return [out[0], out[1], out[2], np.array([out[3], out[4]])]
        idx_p = p.to(dev, non_blocking=True); idx_q = q.to(dev, non_blocking=True)
        idx_dp = dp.to(dev, non_blocking=True) if self.head else None
        idx_dq = dq.to(dev, non_blocking=True) if self.head else None
        lens_host = lens.numpy() if isinstance(lens, torch.Tensor) else np.asarray(lens)
        index = Engine.seq_index(idx_p, idx_q, lens.to(dev, non_blocking=True), idx_dp, idx_dq)
        return self._step(np.arange(B, dtype=np.int32), index, idx_p, idx_q, lens_host)

    def train(self, local_uidx):
        """One step over this rank's users `local_uidx` (int32 indices into the local index matrices).
        Every rank must call it in lock-step.  Returns the GLOBAL [los, sur, upq, ls]."""
        uidx = np.asarray(local_uidx, dtype=np.int32).reshape(-1)
        return self._step(uidx, self._index(), self.P.t, self.Q.t, self._lens_host)

    def _step(self, uidx, index, Pt, Qt, lens_host):
        eng, dev = self.engine, self.engine.torch_device
        B = uidx.size
        if self.peer:
            # one engine call: slice + sort, sharded gather over NVLink, forward + backward, flag exchange, rank-order
            # dense all-reduce, owner-side sparse SGD straight out of the peers' outboxes
            if not self.uniform_batch:
                raise ValueError("the peer-memory step needs the same batch size on every rank (construct with peer=False otherwise)")
            if 2 * B * int(Pt.shape[1]) > self._cap:
                raise ValueError("batch of %d users needs up to %d outbox records, capacity is %d: construct with a larger max_batch"
                                 % (B, 2 * B * int(Pt.shape[1]), self._cap))
            self._step_no += 1
            out = eng.gru_step_mg(self._params(), index, uidx, int(lens_host[uidx].max()), self._peers, self._step_no,
                                  self._alpha, self._lambda)
            return [out[0], out[1], out[2], np.array([out[3], out[4]])]
        if self.trace: self.trace.mark('start')
        meta = torch.tensor([B, int((lens_host[uidx] >= 1).sum())], dtype=torch.int64, device=dev)
        if self.world > 1:
            dist.all_reduce(meta, group=self.group)
        global_batch, n_nonempty = int(meta[0].item()), int(meta[1].item())
        if self.trace: self.trace.mark('meta_allreduce')
        uniq = eng.gru_mg_prepare(self._params(), index, uidx, Pt.shape[1])     # slice + sort once, reused below
        if self.trace: self.trace.mark('prepare(slice+sort)')
        n_u = uniq.numel()
        max_len = int(lens_host[uidx].max())
        ex = RowExchange(uniq, self.world, self.group)
        rows = ex.fetch(lambda loc: eng.gather_rows(self.lt_local.t, loc))
        if self.trace: self.trace.mark('a2a_fetch_rows')
        row_grads = torch.empty((n_u, self.d), dtype=torch.float32, device=dev)
        row_cnt = torch.empty(n_u, dtype=torch.float32, device=dev)
        eng.gru_train_mg(self._params(), index, uidx, max_len, global_batch, rows,
                         self._dense, row_grads, row_cnt, self._sums)
        if self.trace: self.trace.mark('train_mg(fwd+bwd)')
        if self.world > 1:
            dist.all_reduce(self._dense, group=self.group)
            dist.all_reduce(self._sums, group=self.group)
        recv_grads = ex.push(row_grads)
        recv_cnts = ex.push(row_cnt)
        recv_local = ex.recv_local
        self.last_exchange_rows = n_u
        if self.trace: self.trace.mark('exchange_back')
        out = eng.gru_apply_mg(self._params(), self._dense, self._sums, global_batch, 0 if self.head else n_nonempty,
                               self.lt_local.t, recv_local, recv_grads, recv_cnts, self._alpha, self._lambda)
        if self.trace: self.trace.mark('apply')
        if self.trace: self.trace.end_step()
        return [out[0], out[1], out[2], np.array([out[3], out[4]])]

    # ---- checkpoint of the row-sharded model ("sharded save for C5") ----------------------------------
    def save_checkpoint(self, directory, epoch=0):
        """Every rank writes its shard of the item table as `lt.shard<r>of<W>.npy`; rank 0 adds `dense.pkl`: the reference's
        9-array list (prog_bpr_gru_spatial.py:323-330, order fixed by GRU_Spatial.py:92-101) with the `lt` slot holding a
        manifest {n_rows, d, world, pattern} instead of the 20 GB table.  `assemble_checkpoint` turns the directory back into
        one reference-format file when the table fits the host."""
        import os
        import pickle
        os.makedirs(directory, exist_ok=True)
        np.save(os.path.join(directory, "lt.shard%dof%d.npy" % (self.rank, self.world)), self.lt_local.get_value())
        if self.rank == 0:
            sc = self._scal.get_value()
            manifest = dict(n_rows=self.n_rows, d=self.d, world=self.world, pattern="lt.shard%dof%d.npy", epoch=int(epoch))
            arrays = [np.asarray(sc[1:], dtype=np.float32), np.asarray(sc[0], dtype=np.float64), manifest, self.di.get_value(),
                      self.ui.get_value(), self.wh.get_value(), self.bi.get_value(), self.vs.get_value(), self.bs.get_value()]
            with open(os.path.join(directory, "dense.pkl"), "wb") as f:
                pickle.dump(arrays, f, protocol=2)
        if self.world > 1:
            dist.barrier(group=self.group)

    def load_checkpoint(self, directory):
        """Inverse of `save_checkpoint`; the directory may have been written with a different world size (rows are re-dealt:
        global row = local * world_saved + rank_saved)."""
        import os
        import pickle
        with open(os.path.join(directory, "dense.pkl"), "rb") as f:
            arrays = pickle.load(f, encoding="latin1")
        man = arrays[2]
        if man["n_rows"] != self.n_rows or man["d"] != self.d:
            raise ValueError("checkpoint is for a %d x %d table, this model has %d x %d" % (man["n_rows"], man["d"], self.n_rows, self.d))
        mine = np.arange(self.rank, self.n_rows, self.world)
        out = np.empty((len(mine), self.d), dtype=np.float32)
        for r in range(man["world"]):
            sh = np.load(os.path.join(directory, man["pattern"] % (r, man["world"])), mmap_mode="r")
            sel = np.nonzero(mine % man["world"] == r)[0]
            out[sel] = sh[mine[sel] // man["world"]]
        self.lt_local.t.copy_(torch.from_numpy(out))
        self._scal.set_value(np.array([float(arrays[1]), arrays[0][0], arrays[0][1]], dtype=np.float32))
        for k, a in zip(("di", "ui", "wh", "bi", "vs", "bs"), arrays[3:]):
            getattr(self, k).set_value(np.asarray(a, dtype=np.float32))


class ShardedGeoIE:
    """Mini-batch GeoIE with K negatives over `world` GPUs (BASELINE.json C4: "GeoIE ... 2 x B200 row-sharded"; EXTENSION
    semantics as `GeoIEBatch`).  Users are split over the ranks; the item tables g, h, z are ROW-SHARDED (owner = row %
    world), the scalars a, b replicated.  `train_batch(P, Q)` takes THIS rank's users (global POI ids) and runs the
    peer-memory step of csrc/mf_mg.cuh; every rank must call it in lock-step with the same batch shape.  G ranks x Bu
    users is the same update as one GPU x G Bu users up to float32 rounding."""

    def __init__(self, alpha_lambda, n_item, n_hidden, init, coords, max_users, seq_len, n_neg, rank=None, world=None, device=None,
                 tables_are_shards=False, group=None):
        from ._lib import PoiMfPeers
        from .engine import Engine
        from .shared import Shared
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.group = group
        self.engine = Engine.get(device)
        eng, W, dev = self.engine, self.world, self.engine.torch_device
        self.n_rows, self.H = n_item + 1, n_hidden
        self._alpha, self._lambda = float(alpha_lambda[0]), float(alpha_lambda[1])
        self._ab = torch.tensor([float(init["a"]), float(init["b"])], dtype=torch.float64, device=dev)
        from .public.GeoIE import coords_table
        self.coords = Shared(coords_table(coords, self.n_rows), "float32", dev)
        n = seq_len - 1
        cap = (max_users * n, max_users * n * (n_neg + 1))
        own, handles = {}, {}
        for k in ("g", "h", "z"):
            t = init[k]
            if not tables_are_shards:
                t = t[self.rank::W] if isinstance(t, torch.Tensor) else shard_rows(np.asarray(t), self.rank, W)
            t = t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
            own[k], handles[k] = eng.peer_alloc(tuple(t.shape), torch.float32)
            own[k].copy_(t)
        n_local = int(own["g"].shape[0])
        spec = {}
        for s_ in (0, 1):
            spec["ob_ids%d" % s_] = ((cap[s_],), torch.int32); spec["ob_perm%d" % s_] = ((cap[s_],), torch.int32)
            spec["ob_meta%d" % s_] = ((W + 2,), torch.int32)
        for t_, s_ in ((0, 0), (1, 1), (2, 1)):
            spec["ob_grads%d" % t_] = ((cap[s_], n_hidden), torch.float32)
        spec["sums"] = ((4,), torch.float64); spec["flags"] = ((2 * W,), torch.int32)
        for k, (shape, dt) in spec.items():
            own[k], handles[k] = eng.peer_alloc(shape, dt)
            if own[k].numel() <= 4096:
                own[k].zero_()
        self.g, self.h, self.z = (Shared(own[k], "float32", dev) for k in ("g", "h", "z"))
        self._own = own
        self._slot = [torch.full((n_local * W,), -1, dtype=torch.int32, device=dev) for _ in range(2)]
        torch.cuda.synchronize(dev)
        mine = dict(h=handles, shapes={k: tuple(own[k].shape) for k in own}, cap=cap)
        every = [None] * W
        if W > 1:
            dist.all_gather_object(every, mine, group=group)
        else:
            every[0] = mine
        if any(tuple(x["cap"]) != tuple(cap) for x in every):
            raise ValueError("all ranks must use the same max_users / seq_len / n_neg")
        pt = PoiMfPeers()
        pt.world, pt.rank, pt.n_local_rows = W, self.rank, n_local
        pt.cap[0], pt.cap[1] = cap
        self._mapped = []
        dts = {torch.float32: torch.float32, torch.int32: torch.int32, torch.float64: torch.float64}

        def ptr_of(r, x, key):
            if r == self.rank:
                return own[key].data_ptr()
            m = eng.peer_open(x["h"][key], x["shapes"][key], own[key].dtype)
            self._mapped.append(m)
            return m.data_ptr()
        for r, x in enumerate(every):
            for t_, k in enumerate(("g", "h", "z")):
                pt.shard[t_][r] = ptr_of(r, x, k)
                pt.ob_grads[t_][r] = ptr_of(r, x, "ob_grads%d" % t_)
            for s_ in (0, 1):
                pt.ob_ids[s_][r] = ptr_of(r, x, "ob_ids%d" % s_); pt.ob_perm[s_][r] = ptr_of(r, x, "ob_perm%d" % s_)
                pt.ob_meta[s_][r] = ptr_of(r, x, "ob_meta%d" % s_)
            pt.sums[r] = ptr_of(r, x, "sums"); pt.flags[r] = ptr_of(r, x, "flags")
        pt.slot_tab[0], pt.slot_tab[1] = self._slot[0].data_ptr(), self._slot[1].data_ptr()
        self._peers, self._cap, self._step_no = pt, cap, 0
        if W > 1:
            dist.barrier(group=group)

    def train_batch(self, P, Q):
        """P [Bu, L], Q [Bu, L, K] of THIS rank's users (global POI ids; CUDA tensors or host arrays).  Returns the GLOBAL
        summed log-sigmoid loss."""
        Bu, L = int(P.shape[0]), int(P.shape[1]); K = int(Q.shape[2])
        if Bu * (L - 1) > self._cap[0] or Bu * (L - 1) * (K + 1) > self._cap[1]:
            raise ValueError("batch exceeds the outbox capacity: construct with a larger max_users")
        self._step_no += 1
        return self.engine.geoie_step_mg(self._ab, self.n_rows, self.H, P, Q, self.coords.t, self._peers, self._step_no,
                                         self._alpha, self._lambda)

    def a_b(self):
        v = self._ab.cpu().numpy()
        return float(v[0]), float(v[1])


def assemble_checkpoint(directory, out_path):
    """Sharded checkpoint directory -> ONE file in the reference's 9-array format (loadable by `OboSpatialGru.load_params`
    and by the reference itself)."""
    import os
    import pickle
    with open(os.path.join(directory, "dense.pkl"), "rb") as f:
        arrays = pickle.load(f, encoding="latin1")
    man = arrays[2]
    arrays[2] = unshard_rows([np.load(os.path.join(directory, man["pattern"] % (r, man["world"]))) for r in range(man["world"])], man["n_rows"])
    with open(out_path, "wb") as f:
        pickle.dump(arrays, f, protocol=2)
    return out_path
