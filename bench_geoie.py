"""bench.py --config c4: GeoIE, |POI| = 1M, |U| = 100k, seq = 32, d = 256, K = 100 negatives per target (BASELINE.json
configs[3]).  One step = one `GeoIEBatch.train_batch` call over `--batch` users (all 31 targets of each).  HBM-bound:
algorithmic bytes per check-in = 203 rows x 2 x 1 KB + indices = 416 148 B."""
import json
import os
import sys
import time

import numpy as np

import bench as B0

ALGO = 416148.0
# The reference's alpha = 0.01 is a per-user step with ONE negative (31 loss terms per step).  A mini-batch sums, without
# normalisation (Bpr's rule), batch x 31 targets x 100 negatives = 4e5 .. 1e6 terms into the two shared scalars a and b -- and b
# sits in the exponent of a d^b with d up to 50 km: with 0.01, 0.01 / batch or even 2e-6 a few steps send b to ~10 and the
# scores to inf - inf = NaN (measured).  1e-7 keeps 50 steps finite; the arithmetic per check-in does not depend on alpha.
ALPHA_C4 = 1e-7


def _data(cfg, n_users, seed=123):
    """P [n_users x seq], Q [n_users x seq x K] (uniform, outside the user's own sequence), coords fp32 -- only for the users
    the run touches (the full |U| = 100k would be 1.3 GB of negatives)."""
    rs = np.random.RandomState(seed)
    I, L, K = cfg["n_item"], cfg["seq"], cfg["neg"]
    P = rs.randint(0, I, size=(n_users, L)).astype(np.int32)
    Q = rs.randint(0, I, size=(n_users, L, K)).astype(np.int32)
    own = np.sort(P, axis=1)
    for _ in range(4):
        flat = Q.reshape(n_users, -1)
        pos = np.clip(np.stack([np.searchsorted(own[u], flat[u]) for u in range(n_users)]), 0, L - 1)
        hit = (np.take_along_axis(own, pos, axis=1) == flat).reshape(Q.shape)
        if not hit.any():
            break
        Q[hit] = rs.randint(0, I, size=int(hit.sum()))
    coords = np.zeros((I + 1, 2), dtype=np.float32)
    coords[:I, 0] = rs.uniform(1.22, 1.47, I); coords[:I, 1] = rs.uniform(103.60, 104.04, I)
    return P, Q, coords


def cpu_step(OM, state, ds, users, alpha, lam):
    P, Q, coords = ds["P"], ds["Q"], ds["coords"]
    _, state = OM.geoie_train_batch_k(state, users, P[users], Q[users], coords.astype(np.float64), alpha, lam)
    return state, len(users) * (P.shape[1] - 1)


def run_reference(args):
    import poi_b200  # noqa: F401
    import torch
    from oracle import models as OM
    from poi_b200 import synth
    cfg = dict(synth.CONFIGS["c4"])
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cb = 8
    steps, warm = max(1, args.steps), 1
    P, Q, coords = _data(cfg, cb * (steps + warm))
    st = synth.init_mf_state("geoie", 8, cfg["n_item"], cfg["d"])
    state = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    ds = dict(P=P, Q=Q, coords=coords)
    ts, done = [], 0
    for s in range(warm + steps):
        users = np.arange(s * cb, (s + 1) * cb)
        t_ = time.perf_counter()
        state, n_ci = cpu_step(OM, state, ds, users, ALPHA_C4, B0.LAM)
        dt = time.perf_counter() - t_
        if s >= warm:
            ts.append(dt); done += n_ci
    value = done / sum(ts)
    line = {"impl": "reference", "metric": B0.METRIC, "value": value, "unit": B0.UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": sum(ts) / len(ts) * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": B0.make_config("c4", cfg, args.batch, args.gpus, args.scaling),
            "cpu_baseline": {"value": value, "unit": B0.UNIT, "cores": cores, "kind": "port",
                             "sample": "%d steps x %d users of the %d-user step (torch-CPU oracle, float64, all cores)" % (steps, cb, args.batch)},
            "e2e": {"value": value, "unit": B0.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import poi_b200  # noqa: F401
    from poi_b200 import synth
    from poi_b200.public.GeoIE import GeoIEBatch
    cfg = dict(synth.CONFIGS["c4"])
    I, d, L, K = cfg["n_item"], cfg["d"], cfg["seq"], cfg["neg"]
    weak = args.scaling == "weak"
    Bu = args.batch if weak else max(1, args.batch // world)          # users per GPU per step
    Bg = Bu * world
    W, Kst = max(args.warmup, 3), max(args.steps, 1)
    n_steps = 2 * (W + Kst) + 4
    P, Q, coords = _data(cfg, Bg * n_steps)                              # same seed on every rank: identical data
    n_user = Bg * n_steps
    dev = torch.device("cuda", local_rank)
    st = synth.init_mf_state("geoie", 8, I, d)
    tes = [[I]]
    if world == 1:
        model = GeoIEBatch([tes, tes, [[1]], [[1]]], [tes, tes], [ALPHA_C4, B0.LAM], 8, I, d, d, None, init=st, coords=coords, device=local_rank)
    else:
        from poi_b200.dist import ShardedGeoIE
        model = ShardedGeoIE([ALPHA_C4, B0.LAM], I, d, st, coords, max_users=Bu, seq_len=L, n_neg=K, device=local_rank)
    eng = model.engine
    mine = lambda s: slice(s * Bg + rank * Bu, s * Bg + (rank + 1) * Bu)  # this rank's users of step s
    res = [(torch.as_tensor(P[mine(s)], device=dev), torch.as_tensor(Q[mine(s)], device=dev)) for s in range(n_steps)]
    pin = [(torch.from_numpy(P[mine(s)]).pin_memory(), torch.from_numpy(Q[mine(s)]).pin_memory()) for s in range(n_steps)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(arrs, n_warm, n, first):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for i in range(n_warm):
            model.train_batch(*arrs[first + i])
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = eng.launch_count(); losses = []
        for i in range(n):
            flush.fill_(i & 0xff)
            ev[i][0].record()
            losses.append(model.train_batch(*arrs[first + n_warm + i]))
            ev[i][1].record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        return sum(a.elapsed_time(b) for a, b in ev), eng.launch_count() - l0, losses

    sampler = B0.ClockSampler(local_rank); sampler.start()
    ms, launches, losses = timed(res, W, Kst, 0)
    ms_e2e, _, _ = timed(pin, 1, Kst, W + Kst)
    clocks = sampler.stop()
    eng.kprof_reset(); eng.kprof_enable(True)
    nprof = min(Kst, 3)
    for i in range(nprof):
        model.train_batch(*res[2 * (W + Kst) + i])
    prof = eng.kprof_get(); eng.kprof_enable(False)
    if dist is not None:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0].item()), float(t[1].item())
        t = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        launches = int(t.item())
        if rank != 0:
            dist.destroy_process_group()
            return
    peaks = B0.load_peaks()
    N = Bg * (L - 1)
    value = N * Kst / (ms * 1e-3)
    kern = {}
    for k, what in (("geoie", "k_geoie_batch_k: per user, G in shared memory / registers, candidates streamed, unique rows updated in place"),
                    ("rows", "segment sums of the rows that occur several times in the batch (multi-GPU: + deltas, owner-side apply)"),
                    ("gather", "multi-GPU: compact copies of the touched rows out of the owners' shards (NVLink)"),
                    ("index", "keys, radix sort, segments"), ("other", "multi-GPU: flag signal / wait kernels"),
                    ("reduce", "loss / a, b finalisation")):
        r = prof[k]
        if r["ms"] > 0:
            kern[k] = {"what": what, "ms_per_step": r["ms"] / nprof, "launches_per_step": r["launches"] / nprof}
    t_main = (prof["geoie"]["ms"] + prof["rows"]["ms"] + prof["gather"]["ms"]) / nprof
    ach = ALGO * (N / world) / (t_main * 1e-3) / 1e9                     # per GPU
    tot_ms = sum(v["ms"] for v in prof.values()) / nprof
    roof = {"kernel": "k_geoie_batch_k + segment sums (the gather / scatter of the step)", "bound": "hbm", "achieved": ach, "peak": peaks["hbm"],
            "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": None, "peak_source": peaks["source"], "algorithmic_bytes_per_check_in": ALGO,
            "share_of_step": t_main / tot_ms, "whole_step_frac": ALGO * (N / world) / (ms / Kst * 1e-3) / 1e9 / peaks["hbm"]}
    tp = os.path.join(B0.ROOT, "profiles", "r2_ncu_traffic.json")
    if world == 1 and Bg == 128 and os.path.exists(tp):                  # the capture is of the 128-user step
        with open(tp) as f:
            tr = json.load(f).get("kernels", {})
        if "geoie_batch_k" in tr:
            roof["traffic"] = tr["geoie_batch_k"]["dram_bytes"]
    parity = None
    if not args.no_parity and world == 1:
        from oracle import models as OM
        nb = 4
        m2 = GeoIEBatch([tes, tes, [[1]], [[1]]], [tes, tes], [ALPHA_C4, B0.LAM], n_user, I, d, d, None, init=st, coords=coords, device=local_rank)
        got = m2.train_batch(res[0][0][:nb].contiguous(), res[0][1][:nb].contiguous())
        ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
        want, ref = OM.geoie_train_batch_k(ref, np.arange(nb), P[:nb], Q[:nb], coords.astype(np.float64), ALPHA_C4, B0.LAM)
        rows = np.unique(np.concatenate((P[:nb].ravel(), Q[:nb, 1:].ravel())))

        def el(a, b):
            a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
            return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-2 * np.max(np.abs(b)))))
        parity = {"users": nb, "oracle": "oracle.models.geoie_train_batch_k float64", "rel_err_loss": abs(got - want) / abs(want),
                  "rel_err_rows": max(el(getattr(m2, k).get_value()[rows], ref[k][rows]) for k in ("g", "h", "z")),
                  "rel_err_ab": max(abs(m2.a.eval() - ref["a"]) / abs(ref["a"]), abs(m2.b.eval() - ref["b"]) / abs(ref["b"])), "tolerance": 1e-4,
                  "metric": "element-wise |a-b| / max(|b|, 1e-2 max|b|) on the touched rows"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import models as OM
        torch.set_num_threads(os.cpu_count() or 1)
        state = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
        ds = dict(P=P, Q=Q, coords=coords)
        ts = []
        for s in range(3):
            t_ = time.perf_counter()
            state, n_ci = cpu_step(OM, state, ds, np.arange(s * 8, s * 8 + 8), ALPHA_C4, B0.LAM)
            ts.append(time.perf_counter() - t_)
        cpu = {"value": n_ci * 2 / sum(ts[1:]), "unit": B0.UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "2 steps x 8 users of the %d-user step (torch-CPU oracle, float64, all cores)" % Bu}
    if not np.isfinite(losses[-1]):
        print("bench c4: WARNING the loss is not finite (%r): a / b have diverged" % losses[-1], file=sys.stderr, flush=True)
    line = {"metric": B0.METRIC, "value": value, "unit": B0.UNIT, "n_gpus": world, "steps": Kst, "warmup": W, "ms_per_step": ms / Kst,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": B0.make_config("c4", cfg, args.batch, world, args.scaling),
            "engine": {"check_ins_per_step": N, "negatives": K, "users_per_gpu": Bu,
                       "parallelism": "1 GPU holding all three 1 GB tables" if world == 1 else
                       "dp%d: users sharded, g / h / z row-sharded (row %% %d), compact copies gathered from the owners' shards and deltas applied by the "
                       "owners over NVLink peer memory (csrc/mf_mg.cuh)" % (world, world)},
            "e2e": {"value": N * Kst / (ms_e2e * 1e-3), "unit": B0.UNIT, "h2d_bytes_per_step": Bu * L * (K + 1) * 4, "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e / Kst},
            "gpu_launches": int(launches), "roofline": roof, "kernels": kern,
            "kernel_ms_per_step": {k: round(v["ms"] / nprof, 4) for k, v in prof.items() if v["ms"] > 0},
            "parity": parity, "cpu_baseline": cpu, "clocks": clocks, "final_loss": float(losses[-1])}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
