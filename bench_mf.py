"""bench.py --config c3 | c4: the pairwise embedding models with K negatives per positive.

  c3  PRME   |POI| = 100k, |U| = 10k, seq = 32, d = 256, K = 20   (BASELINE.json configs[2], 1 x B200)
  c4  GeoIE  |POI| = 1M,  |U| = 100k, seq = 32, d = 256, K = 100  (BASELINE.json configs[3], 2 x B200 row-sharded)

One step = one mini-batch call (throughput mode, EXTENSION semantics: the reference trains these models one check-in / one
user at a time with one negative; the mini-batch rule is the reference's own Bpr one, BPR.py:351-397) over `--batch`
users x `--positions` consecutive positions.  These paths are HBM-bound: `roofline.bound` = "hbm", algorithmic bytes per
check-in from SURVEY.md 8d (c3: 92 260 B, c4: 416 148 B)."""
import json
import os
import sys
import time

import numpy as np

import bench as B0

ALGO_BYTES = {"c3": 92260.0, "c4": 416148.0}


def _workload(cfg_name, n_user_cap=None):
    import poi_b200  # noqa: F401
    from poi_b200 import synth
    cfg = dict(synth.CONFIGS[cfg_name])
    nu = cfg["n_user"] if n_user_cap is None else min(cfg["n_user"], n_user_cap)
    ds = synth.make_mf_dataset(nu, cfg["n_item"], cfg["seq"], cfg["neg"])
    st = synth.init_mf_state(cfg["model"], nu, cfg["n_item"], cfg["d"])
    return cfg, ds, st


def prme_step_arrays(ds, users, t0, npos):
    """The (u, p, Q, prev, dist, gap) arrays of one step: positions t0 .. t0+npos-1 (1-based targets) of `users`."""
    ts = np.arange(t0, t0 + npos)
    uu = np.repeat(users, npos); tt = np.tile(ts, len(users))
    return (uu.astype(np.int32), ds["P"][uu, tt], np.ascontiguousarray(ds["Q"][uu, tt]), ds["P"][uu, tt - 1],
            ds["dist"][uu, tt].astype(np.float32), ds["gap"][uu, tt])


def _plan(ds, batch, npos, n_steps):
    """Step s -> (user block, first position): user blocks first, then the next group of positions."""
    U, T = ds["n_user"], ds["seq"] - 1
    nb = max(1, U // batch)
    groups = max(1, T // npos)
    out = []
    for s in range(n_steps):
        blk, grp = s % nb, (s // nb) % groups
        out.append((np.arange(blk * batch, blk * batch + batch) % U, 1 + grp * npos))
    return out


def run_reference(args):
    if args.config == "c4":
        import bench_geoie
        return bench_geoie.run_reference(args)
    cfg, ds, st = _workload(args.config)
    from oracle import models as OM
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    npos = args.positions
    cb = max(1, min(args.cpu_batch, ds["n_user"]) // (4 if args.config == "c3" else 16))
    steps, warm = max(1, args.steps), 1
    state = {k: np.asarray(v, dtype=np.float32) for k, v in st.items()}
    times, done = [], 0
    for s, (users, t0) in enumerate(_plan(ds, cb, npos, warm + steps)):
        t_ = time.perf_counter()
        if args.config == "c3":
            a = prme_step_arrays(ds, users, t0, npos)
            _, state = OM.prme_train_batch_k(state, a[0], a[1], a[2], a[3], a[4], a[5], B0.ALPHA, B0.LAM, 360, 0.2, dtype=torch.float32)
            n_ci = len(a[0])
        else:
            import bench_geoie
            state, n_ci = bench_geoie.cpu_step(OM, state, ds, users, B0.ALPHA, B0.LAM)
        dt = time.perf_counter() - t_
        if s >= warm:
            times.append(dt); done += n_ci
    value = done / sum(times)
    line = {"impl": "reference", "metric": B0.METRIC, "value": value, "unit": B0.UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": sum(times) / len(times) * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": B0.make_config(args.config, cfg, args.batch, args.gpus, args.scaling),
            "cpu_baseline": {"value": value, "unit": B0.UNIT, "cores": cores, "kind": "port",
                             "sample": "%d steps x %d users of the %d-user step (torch-CPU oracle, float32, all cores)" % (steps, cb, args.batch)},
            "e2e": {"value": value, "unit": B0.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config == "c4":
        import bench_geoie
        return bench_geoie.run_ours(args)
    if world > 1:
        raise SystemExit("bench.py --config c3 is a single-GPU workload (BASELINE.json configs[2]: 1 x B200)")
    import poi_b200  # noqa: F401
    from poi_b200.public.PRME import Prme
    cfg, ds, st = _workload("c3")
    U, I, d, K = ds["n_user"], ds["n_item"], cfg["d"], cfg["neg"]
    dev = torch.device("cuda", local_rank)
    tes = [[I]]
    side = [tes, [[0]], [[0.0]], [[1]], tes]
    model = Prme(side, side, [B0.ALPHA, B0.LAM], 360, 0.2, ds["coords"], U, I, d, init=st, device=local_rank)
    eng = model.engine
    Bu, npos = min(args.batch, U), args.positions
    W, Kst = max(args.warmup, 3), max(args.steps, 1)
    plan = _plan(ds, Bu, npos, 2 * (W + Kst) + 8)
    host = [prme_step_arrays(ds, u, t0, npos) for (u, t0) in plan]
    N = len(host[0][0])
    dts = (torch.int32, torch.int32, torch.int32, torch.int32, torch.float32, torch.int32)
    resident = [tuple(torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev) for a, dt in zip(h, dts)) for h in host]
    pinned = [tuple(torch.as_tensor(np.ascontiguousarray(a), dtype=dt).pin_memory() for a, dt in zip(h, dts)) for h in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(arrs, n_warm, n_steps, first):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        for i in range(n_warm):
            model.train(*arrs[first + i])
        torch.cuda.synchronize()
        l0 = eng.launch_count(); losses = []
        for i in range(n_steps):
            flush.fill_(i & 0xff)
            ev[i][0].record()
            losses.append(model.train(*arrs[first + n_warm + i]))
            ev[i][1].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev), eng.launch_count() - l0, losses

    sampler = B0.ClockSampler(local_rank); sampler.start()
    ms, launches, losses = timed(resident, W, Kst, 0)
    ms_e2e, _, _ = timed(pinned, 1, Kst, W + Kst)
    clocks = sampler.stop()
    eng.kprof_reset(); eng.kprof_enable(True)
    nprof = min(Kst, 3)
    for i in range(nprof):
        model.train(*resident[2 * (W + Kst) + i])
    prof = eng.kprof_get(); eng.kprof_enable(False)
    peaks = B0.load_peaks()
    value = N * Kst / (ms * 1e-3)
    algo = ALGO_BYTES["c3"] * N
    kern = {}
    for k, what in (("mf", "k_prme_score: gather 2(K+2)+1 rows per check-in, distances, loss, per-occurrence scalars"),
                    ("rows", "k_prme_apply + user rows: one read-modify-write per unique row of the batch"),
                    ("index", "keys, radix sort, segments"), ("reduce", "loss partials")):
        r = prof[k]
        if r["ms"] > 0:
            kern[k] = {"what": what, "ms_per_step": r["ms"] / nprof, "launches_per_step": r["launches"] / nprof}
            if r["bytes"] > 0:
                a = r["bytes"] / (r["ms"] * 1e-3) / 1e9
                kern[k].update(achieved=a, unit="GB/s", peak=peaks["hbm"], frac=a / peaks["hbm"], algorithmic_bytes_per_step=r["bytes"] / nprof)
    t_row = (prof["mf"]["ms"] + prof["rows"]["ms"]) / nprof
    ach = algo / (t_row * 1e-3) / 1e9
    tot_ms = sum(v["ms"] for v in prof.values()) / nprof
    roof = {"kernel": "k_prme_score + k_prme_apply (the gather / scatter pair of the step)", "bound": "hbm", "achieved": ach, "peak": peaks["hbm"],
            "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": None, "peak_source": peaks["source"],
            "algorithmic_bytes_per_check_in": ALGO_BYTES["c3"], "share_of_step": t_row / tot_ms,
            "whole_step_frac": algo / (ms / Kst * 1e-3) / 1e9 / peaks["hbm"]}
    tp = os.path.join(B0.ROOT, "profiles", "r2_ncu_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tr = json.load(f).get("kernels", {})
        if "prme_score" in tr and "prme_apply" in tr:
            roof["traffic"] = tr["prme_score"]["dram_bytes"] + tr["prme_apply"]["dram_bytes"]
    # parity: one shared step on both arms, fresh models
    parity = None
    if not args.no_parity:
        from oracle import models as OM
        n1 = min(1024, N)
        m2 = Prme(side, side, [B0.ALPHA, B0.LAM], 360, 0.2, ds["coords"], U, I, d, init=st, device=local_rank)
        got = m2.train(*[a[:n1].contiguous() for a in resident[0]])
        h = [a[:n1] for a in host[0]]
        ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
        want, ref = OM.prme_train_batch_k(ref, h[0], h[1], h[2], h[3], h[4], h[5], B0.ALPHA, B0.LAM, 360, 0.2)
        rows = np.unique(np.concatenate((h[1], h[3], h[2].ravel())))

        def el(a, b):
            a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
            return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-2 * np.max(np.abs(b)))))
        parity = {"check_ins": int(n1), "oracle": "oracle.models.prme_train_batch_k float64", "rel_err_loss": abs(got - want) / abs(want),
                  "rel_err_rows": max(el(m2.dp.get_value()[rows], ref["dp"][rows]), el(m2.ds.get_value()[rows], ref["ds"][rows]),
                                      el(m2.du.get_value()[h[0]], ref["du"][h[0]])), "tolerance": 1e-4, "metric": "element-wise |a-b| / max(|b|, 1e-2 max|b|) on the touched rows (float32 cancellation in D(q)-D(p) at d=256, see tests/test_gpu_prme_k.py)"}
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import models as OM
        torch.set_num_threads(os.cpu_count() or 1)
        state = {k: np.asarray(v, dtype=np.float32) for k, v in st.items()}
        n1 = min(1024, N)
        ts = []
        for s in range(3):
            h = [a[:n1] for a in host[s]]
            t_ = time.perf_counter()
            _, state = OM.prme_train_batch_k(state, h[0], h[1], h[2], h[3], h[4], h[5], B0.ALPHA, B0.LAM, 360, 0.2, dtype=torch.float32)
            ts.append(time.perf_counter() - t_)
        cpu = {"value": n1 * 2 / sum(ts[1:]), "unit": B0.UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "2 steps x %d check-ins of the %d-check-in step (torch-CPU oracle, float32, all cores)" % (n1, N)}
    h2d = N * (K + 5) * 4
    line = {"metric": B0.METRIC, "value": value, "unit": B0.UNIT, "n_gpus": 1, "steps": Kst, "warmup": W, "ms_per_step": ms / Kst,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": B0.make_config("c3", cfg, args.batch, 1, args.scaling),
            "engine": {"check_ins_per_step": N, "positions_per_step": npos, "negatives": K,
                       "path": "k_prme_score (scalars per occurrence) + k_prme_apply (one RMW per unique row); fp32 FMA, no tensor cores"},
            "e2e": {"value": N * Kst / (ms_e2e * 1e-3), "unit": B0.UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / Kst},
            "gpu_launches": int(launches), "roofline": roof, "kernels": kern,
            "kernel_ms_per_step": {k: round(v["ms"] / nprof, 4) for k, v in prof.items() if v["ms"] > 0},
            "parity": parity, "cpu_baseline": cpu, "clocks": clocks, "final_loss": float(losses[-1])}
    print(json.dumps(line), flush=True)
