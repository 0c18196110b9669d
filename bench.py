#!/usr/bin/env python
"""bench.py -- train check-ins/sec of the next-POI hot path on synthetic check-in sequences.

Default workload (BASELINE.json configs[1], "c2"): Distance2Pre, |POI| = 40k, |U| = 10k, seq = 32, d = H = 128, 201
distance intervals, alpha = 0.01, lambda = 0.001; one step = one mini-batch `SpatialGru.train` call over `--batch` users
per GPU (gather -> GRU recurrence -> interval-softmax head -> BPR + survival loss -> BPTT -> dense SGD -> sparse row SGD).
A check-in = one valid training position (L-1 per user).  Other workloads: `--config c3` (PRME, K = 20 negatives),
`--config c4` (GeoIE, K = 100 negatives), `--config c5` (Distance2Pre |POI| = 10M, d = 512, seq = 256; `--scaling strong`
keeps the GLOBAL batch fixed as N grows).

  value      device-timed (CUDA events on the engine stream), index matrices resident in HBM
  e2e        the same step through the host-rows entry: the batch's index rows come from pinned host memory every step
             (H2D inside the timed region) and the loss scalars are read back
  roofline   the dominant kernel of the step, from the engine's per-launch CUDA-event profiler, against
             MEASURED_PEAKS.json; `kernels` carries the same figures for every kernel group of the step
  parity     one shared 512-user step on both arms (fresh models): engine vs float64 oracle, element-wise
  cpu_baseline / --impl reference
             the CPU oracle (numpy restatement of the reference semantics, all host threads via BLAS) on a bounded
             sample of the same workload; `cpu_baseline_obo` = the reference's own one-by-one semantics incl. the dense
             (I+1) x d gradient (GRU_Spatial.py:212-215)

python bench.py --gpus N --steps K --warmup W [--impl reference] [--batch B] [--config c2|c3|c4|c5] [--scaling weak|strong]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "train check-ins/sec"
UNIT = "check-ins/s"
ALPHA, LAM = 0.01, 0.001
PARITY_USERS = 512


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


def build_workload(cfg_name, users_cap=None, user_mult=1):
    import poi_b200  # noqa: F401
    from poi_b200 import synth
    cfg = dict(synth.CONFIGS[cfg_name])
    # weak scaling: the user population grows with the number of GPUs (each rank owns cfg["n_user"] users)
    n_user = (cfg["n_user"] if users_cap is None else min(cfg["n_user"], users_cap)) * user_mult
    ds = synth.make_dataset(n_user, cfg["n_item"], cfg["seq"], UD=cfg["UD"], dd=cfg["dd"])
    st = synth.init_state(cfg["n_item"], cfg["d"], cfg["d"], ds["dist_num"])
    return cfg, ds, st


def checkins_of(lens):
    return int(np.maximum(np.asarray(lens, dtype=np.int64) - 1, 0).sum())


def make_config(cfg_name, cfg, batch, world, scaling):
    """The `config` object both arms print (same keys, same strings): names the workload, not the implementation."""
    names = {"distance2pre": "Distance2Pre", "prme": "PRME", "geoie": "GeoIE", "gru": "GRU"}
    w = "%s: %s |POI|=%d |U|=%d seq=%d d=%d" % (cfg_name, names[cfg["model"]], cfg["n_item"], cfg["n_user"], cfg["seq"], cfg["d"])
    if cfg["model"] == "distance2pre":
        w += " D=%d" % int(cfg["UD"] * 1000 / cfg["dd"])
    if cfg.get("neg"):
        w += " neg=%d" % cfg["neg"]
    per_gpu = batch if scaling == "weak" else max(1, batch // world)
    return {"workload": w, "users_per_step_per_gpu": per_gpu, "global_users_per_step": per_gpu * world, "scaling": scaling,
            "semantics": "mini-batch extension (SURVEY 3.6); B=1 is the reference's one-by-one mode",
            "l2": "flushed between timed steps (256 MB write)", "n_gpus": world}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy explicit restatement), mini-batch semantics identical to the GPU arm
# --------------------------------------------------------------------------------------------------
def cpu_steps(ds, st, batch, n_steps, warmup, dtype=np.float32):
    from oracle import explicit as E
    ref = {k: np.asarray(v, dtype=dtype) for k, v in st.items()}
    U = ds["n_user"]
    times, done = [], 0
    for i in range(warmup + n_steps):
        s = (i * batch) % U
        se = np.arange(s, min(s + batch, U))
        t0 = time.perf_counter()
        _, ref = E.gru_family_train_batch(ref, ds["P"][se], ds["Q"][se], ds["M"][se], ALPHA, LAM,
                                          ds["DP"][se], ds["DQ"][se], dtype=dtype)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt); done += checkins_of(ds["lens"][se])
    return done / sum(times), sum(times) / len(times)


def cpu_obo(ds, st, n_users, budget_s=20.0):
    """The reference's OWN update semantics on the CPU: one `seq_train(uidx)` per user, including the dense (I+1) x d
    gradient `T.grad(cost, self.lt)` materialises per call (GRU_Spatial.py:212-215) -- oracle.models with dense=True."""
    import torch
    from oracle import models as OM
    state = {k: np.asarray(v, dtype=np.float32) for k, v in st.items()}
    done, t_tot, n = 0, 0.0, 0
    for u in range(n_users):
        t0 = time.perf_counter()
        _, state = OM.obo_spatial_gru_train(state, ds["P"][u], ds["Q"][u], ds["DP"][u], ds["DQ"][u], ds["M"][u], ALPHA, LAM,
                                            dtype=torch.float32, dense=True)
        t_tot += time.perf_counter() - t0
        done += checkins_of(ds["lens"][u:u + 1]); n += 1
        if t_tot > budget_s:
            break
    return {"value": done / t_tot, "unit": UNIT, "ms_per_user_call": t_tot / n * 1e3, "users": n, "kind": "port",
            "cores": os.cpu_count() or 1,
            "note": "one-by-one (reference semantics) on the CPU oracle, float32, dense (I+1) x d item-table gradient per "
                    "call as T.grad(cost, self.lt) builds it (GRU_Spatial.py:212-215)"}


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle port; Theano is not installable,
    SURVEY.md 8c) on the box's host cores, same config / metric / unit, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config in ("c3", "c4"):
        import bench_mf
        return bench_mf.run_reference(args)
    cfg, ds, st = build_workload(args.config) if args.config != "c5" else _c5_cpu_workload()
    cores = os.cpu_count() or 1
    try:
        import torch
        torch.set_num_threads(cores)
    except Exception:
        pass
    batch = min(args.cpu_batch, ds["n_user"])
    steps = max(1, args.steps)
    value, sec = cpu_steps(ds, st, batch, steps, max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.config, cfg, args.batch, args.gpus, args.scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d steps x %d users of the %d-user step (numpy oracle, float32, BLAS threads = all cores; "
                                   "cost per check-in is batch-independent on the CPU)" % (steps, batch, args.batch)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _c5_cpu_workload():
    """CPU sample of c5: the same shapes per check-in (d = H = 512, seq = 256, 201 intervals) on a catalogue cut to what the
    host can hold in seconds (1M POIs; the per-check-in cost of the mini-batch oracle does not depend on the table size)."""
    import poi_b200  # noqa: F401
    from poi_b200 import synth
    cfg = dict(synth.CONFIGS["c5"])
    ds = synth.make_dataset(64, 1000000, cfg["seq"], UD=cfg["UD"], dd=cfg["dd"])
    st = synth.init_state(1000000, cfg["d"], cfg["d"], ds["dist_num"])
    for k in ("ui", "wh", "vs"):
        st[k] = (st[k] * (4.0 / np.sqrt(cfg["d"]))).astype(np.float32)
    return cfg, ds, st


def hbm_microbench(eng, dev, peaks, rows=1000001, d=256, n=1 << 20, reps=10):
    """Stand-alone embedding gather and sparse-SGD scatter on a table larger than L2 (C4 shape: 1M x 256 fp32 =
    1.02 GB), uniform random rows -- the "embedding gather HBM GB/s vs roofline" half of BASELINE.json's metric.
    Timed per launch with CUDA events on the engine stream (the engine's own launch profiler).  Algorithmic bytes
    (SURVEY.md 8d): gather = n*d*4 read + the same written + indices; scatter = each touched table row read and
    written once + one gradient row per occurrence + indices."""
    import torch
    g = torch.Generator(device=dev); g.manual_seed(123)
    table = torch.empty((rows, d), dtype=torch.float32, device=dev).uniform_(-0.5, 0.5, generator=g)
    idx = torch.randint(0, rows - 1, (n,), dtype=torch.int32, device=dev, generator=g)
    out = torch.empty((n, d), dtype=torch.float32, device=dev)
    res = {}
    for name in ("gather", "scatter_sgd"):
        def run():
            if name == "gather":
                eng.gather_rows(table, idx, out)
            else:
                eng.scatter_sgd(table, idx, out, ALPHA * 1e-3, LAM)
        for _ in range(3):
            run()
        eng.kprof_reset(); eng.kprof_enable(True)
        for _ in range(reps):
            run()
        prof = eng.kprof_get(); eng.kprof_enable(False)
        if name == "gather":
            ms = prof["gather"]["ms"] / reps
            by = 2.0 * n * d * 4 + 4.0 * n
            parts = {"gather": ms}
        else:
            n_unique = int(torch.unique(idx).numel())
            by = 2.0 * n_unique * d * 4 + 1.0 * n * d * 4 + 4.0 * n
            parts = {k: v["ms"] / reps for k, v in prof.items() if v["ms"] > 0}
            ms = sum(parts.values())
        gbs = by / (ms * 1e-3) / 1e9
        res[name] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                     "ms": ms, "algorithmic_bytes": by, "kernel_ms": {k: round(v, 4) for k, v in parts.items()},
                     "workload": "table %d x %d fp32 (%.2f GB), %d uniform random rows" % (rows, d, rows * d * 4 / 1e9, n)}
    # the row-update kernels alone (the grouping of duplicate rows is integer work on 4-byte keys, not table traffic)
    sc = res["scatter_sgd"]
    if sc["kernel_ms"].get("rows"):
        g = sc["algorithmic_bytes"] / (sc["kernel_ms"]["rows"] * 1e-3) / 1e9
        sc["rows_kernels_only"] = {"achieved": g, "frac": g / peaks["hbm"], "ms": sc["kernel_ms"]["rows"]}
    del table, out, idx
    torch.cuda.empty_cache()
    return res


def parity_block(ds, st, cfg, dev_index, n_users=PARITY_USERS):
    """One shared step on both arms, fresh models from the same initial arrays: the engine (default settings) against the
    float64 oracle.  rel_err_loss = max over the three loss scalars; rel_err_rows / rel_err_dense = element-wise
    |a - b| / max(|b|, 1e-3 max|b|) over the touched item rows / the dense weights (di, ui, wh, vs); the zero-initialised
    biases relative to their largest entry."""
    from oracle import explicit as E
    from poi_b200.public.GRU_Spatial import SpatialGru
    U, I, d, D = ds["n_user"], ds["n_item"], cfg["d"], ds["dist_num"]
    n = min(n_users, U)
    tes = ds["tes"]
    m = SpatialGru([ds["P"], ds["M"], ds["Q"]], [tes, np.ones_like(tes), tes], [ds["DP"], np.full_like(tes, D), ds["DQ"]],
                   [ALPHA, LAM], U, I, [D, cfg["dd"] / 1000.0], d, d, init=st, device=dev_index)
    se = np.arange(n, dtype=np.int32)
    out = m.train(se)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in st.items()}
    (rl, rs, ru, _), ref = E.gru_family_train_batch(ref, ds["P"][se], ds["Q"][se], ds["M"][se], ALPHA, LAM, ds["DP"][se], ds["DQ"][se])

    def el(a, b):
        a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b)) + 1e-300)))
    touched = np.unique(np.concatenate((ds["P"][se].ravel(), ds["Q"][se].ravel())))
    dense = max(el(getattr(m, k).get_value(), ref[k]) for k in ("di", "ui", "wh", "vs"))
    # bi / bs start at zero: after one step they ARE -alpha x a gradient sum over every (t, b) with cancelling signs, so their
    # small entries carry the fp32 summation noise of the large ones -> measured against the largest entry (tests/util.py floor 1)
    bias = max(float(np.max(np.abs(np.asarray(getattr(m, k).get_value(), dtype=np.float64) - ref[k])) / np.max(np.abs(ref[k])))
               for k in ("bi", "bs"))
    return {"users": int(n), "oracle": "oracle.explicit.gru_family_train_batch float64",
            "rel_err_loss": max(abs(a - b) / abs(b) for a, b in zip(out[:3], (rl, rs, ru))),
            "rel_err_rows": el(m.lt.get_value()[touched], ref["lt"][touched]), "rel_err_dense": dense,
            "rel_err_zero_init_biases": bias,
            "loss": float(out[0]), "oracle_loss": float(rl), "tolerance": 1e-4}


# --------------------------------------------------------------------------------------------------
# GPU arm (GRU family: c2, c5)
# --------------------------------------------------------------------------------------------------
KERNEL_GROUPS = {
    # profiler category -> (what it is, bound)
    "gemm": ("tcgen05 GEMMs over all (t, b): input projection, head logits, Vs^T dO, DA.Ui", "tensor"),
    "recur_fwd": ("forward recurrence (fused persistent kernel, or two GEMMs per step when H > 128)", "tensor"),
    "recur_bwd": ("backward recurrence (BPTT through the cell)", "tensor"),
    "wgrad": ("weight-gradient GEMMs + split reduction + dense SGD", "tensor"),
    "gather": ("embedding gather into the time-major input tiles", "hbm"),
    "rows": ("sparse row SGD (segment gather-reduce)", "hbm"),
    "loss": ("interval softmax + survival / BPR loss head", "hbm"),
    "index": ("index slicing, sort, unique, segments", "latency"),
    "eltwise": ("small element-wise kernels (operand staging, transposes)", "hbm"),
    "reduce": ("loss / scalar finalisation", "latency"),
}


def kernel_table(prof, nprof, peaks, gemm_mode):
    tot = sum(v["ms"] for v in prof.values()) or 1.0
    out = {}
    for k, (what, bound) in KERNEL_GROUPS.items():
        r = prof.get(k)
        if not r or r["ms"] <= 0:
            continue
        row = {"what": what, "ms_per_step": r["ms"] / nprof, "share_of_step": r["ms"] / tot, "launches_per_step": r["launches"] / nprof,
               "bound": bound}
        if bound == "tensor" and r["flops"] > 0:
            a = r["flops"] / (r["ms"] * 1e-3) / 1e12
            row.update(achieved=a, unit="TFLOP/s", peak=peaks["tf_sust"], frac=a / peaks["tf_sust"])
            if gemm_mode == 1:
                row["frac_of_3xtf32_ceiling"] = a / (peaks["tf_sust"] / 6.0)
        elif bound == "hbm" and r["bytes"] > 0:
            a = r["bytes"] / (r["ms"] * 1e-3) / 1e9
            row.update(achieved_algorithmic=a, unit="GB/s", peak=peaks["hbm"])
            # no `frac` when the working set is L2-resident (the bytes never reach HBM): see hbm_microbench for the HBM figure
        out[k] = row
    return out


def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path is CUDA only (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if args.config in ("c3", "c4"):
        import bench_mf
        return bench_mf.run_ours(args)
    if args.config == "c5":
        import bench_c5
        return bench_c5.run_ours(args)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import poi_b200  # noqa: F401
    from poi_b200.public.GRU_Spatial import SpatialGru

    weak = args.scaling == "weak"
    cfg, ds, st = build_workload(args.config, user_mult=world if weak else 1)
    U, I, d, seq, D = ds["n_user"], ds["n_item"], cfg["d"], ds["seq"], ds["dist_num"]
    tes = ds["tes"]
    dev = torch.device("cuda", local_rank)
    B_cfg = args.batch if weak else max(1, args.batch // world)
    if world == 1:
        model = SpatialGru([ds["P"], ds["M"], ds["Q"]], [tes, np.ones_like(tes), tes],
                           [ds["DP"], np.full_like(tes, D), ds["DQ"]], [ALPHA, LAM], U, I, [D, cfg["dd"] / 1000.0],
                           d, d, init=st, device=local_rank)
        mine = np.arange(U)
    else:
        # users sharded over ranks, item table row-sharded (owner = row % world), dense weights replicated
        from poi_b200.dist import ShardedSpatialGru
        mine = np.arange(rank, U, world)
        model = ShardedSpatialGru([ds["P"][mine], ds["M"][mine], ds["Q"][mine]], [ds["DP"][mine], ds["DQ"][mine]],
                                  [ALPHA, LAM], I, D, d, d, st, device=local_rank, peer=bool(args.peer),
                                  max_batch=min(B_cfg, len(mine)))
    eng = model.engine
    eng.set_gemm_mode(args.gemm_mode)
    eng.set_fused_recurrence(bool(args.fused))
    eng.set_fused_cluster(args.fused_cluster)
    U_loc = len(mine)
    B = min(B_cfg, U_loc)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def batch_users(i):
        s = (i * B) % U_loc
        return (np.arange(s, s + B) % U_loc).astype(np.int32)

    pinned = {}
    for k in ("P", "Q", "DP", "DQ"):
        pinned[k] = torch.from_numpy(np.ascontiguousarray(ds[k][mine])).pin_memory()
    lens_pin = torch.from_numpy(np.ascontiguousarray(ds["lens"][mine])).pin_memory()
    lens_loc = ds["lens"][mine]

    def step_resident(i):
        return model.train(batch_users(i))

    # pinned staging for the batch's index rows (allocated once; the per-step host work is the row gather into it)
    stage = {k: torch.empty((B, pinned[k].shape[1]), dtype=pinned[k].dtype).pin_memory() for k in pinned}
    stage_len = torch.empty((B,), dtype=lens_pin.dtype).pin_memory()

    def step_host_rows(i):
        se = torch.from_numpy(batch_users(i).astype(np.int64))
        for k in ("P", "Q", "DP", "DQ"):
            torch.index_select(pinned[k], 0, se, out=stage[k])
        torch.index_select(lens_pin, 0, se, out=stage_len)
        return model.train_host_rows(stage["P"], stage["Q"], stage["DP"], stage["DQ"], stage_len)

    def timed(step_fn, n_warm, n_steps, first_step, do_flush=True):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        for i in range(n_warm):
            step_fn(first_step + i)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        if getattr(model, "trace", None) is not None:
            model.trace.acc, model.trace.n = {}, 0         # phase trace: timed steps only
        done = 0
        losses = []
        launches0 = eng.launch_count()
        wall0 = time.perf_counter()
        for i in range(n_steps):
            if do_flush:
                flush.fill_(i & 0xff)                   # evict L2 between timed iterations (outside the events)
            ev[i][0].record()
            out = step_fn(first_step + n_warm + i)
            ev[i][1].record()
            done += checkins_of(lens_loc[batch_users(first_step + n_warm + i)])
            losses.append(out[0])
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        if dist is not None:
            dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        return ms, done, losses, eng.launch_count() - launches0, wall

    W, K = max(args.warmup, 3), max(args.steps, 1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, done, losses, launches, _ = timed(step_resident, W, K, 0)       # launches: my kernels inside the K timed steps
    ms_e2e, done_e2e, _, _, _ = timed(step_host_rows, 1, K, W + K)
    # sustained figure: back-to-back steps for >= args.sustain_s seconds of wall clock (no L2 flush in between; host
    # overhead included) -- the long-run number next to the K-step timed region
    # (the step count must be the SAME on every rank -- the ranks run in lock-step -- so it is derived from the slowest
    # rank's time)
    ms_ref = ms
    if dist is not None:
        t_ = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        ms_ref = float(t_.item())
    k_sus = max(K, int(args.sustain_s / max(ms_ref / K * 1e-3, 1e-5)) + 1) if args.sustain_s > 0 else 0
    sus = None
    if k_sus:
        ms_s, done_s, _, _, wall_s = timed(step_resident, 0, k_sus, 2 * (W + K), do_flush=False)
        sus = (ms_s, done_s, wall_s, k_sus)
    clocks = sampler.stop()

    # per-launch profile of the same steps (CUDA events around every kernel on the engine stream)
    eng.kprof_reset(); eng.kprof_enable(True)
    nprof = min(K, 3)
    for i in range(nprof):
        step_resident(3 * (W + K) + i)
    prof = eng.kprof_get()
    eng.kprof_enable(False)

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if getattr(model, "trace", None) is not None and rank == 0:
        print("[mg phase trace, ms per step, rank 0]", json.dumps(model.trace.report()), file=sys.stderr, flush=True)
        print("[engine kernel ms per profiled step]", json.dumps({k: round(v["ms"] / nprof, 4) for k, v in prof.items() if v["ms"] > 0}),
              file=sys.stderr, flush=True)
    ms_max, ms_e2e_max = allmax(ms), allmax(ms_e2e)
    done_all, done_e2e_all = allsum(done), allsum(done_e2e)
    launches_all = int(allsum(launches))
    sustained = None
    if sus:
        sustained = {"value": allsum(sus[1]) / allmax(sus[2]), "unit": UNIT, "steps": sus[3], "wall_s": allmax(sus[2]),
                     "device_ms_per_step": allmax(sus[0]) / sus[3],
                     "note": "back-to-back steps, wall clock incl. host overhead, no L2 flush between steps"}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    value = done_all / (ms_max * 1e-3)
    e2e_val = done_e2e_all / (ms_e2e_max * 1e-3)
    kernels = kernel_table(prof, nprof, peaks, args.gemm_mode)
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("kernels", {})
    # dominant kernel of the step = the group with the largest share among the tensor / hbm bound ones
    dom = max((k for k in kernels if kernels[k]["bound"] in ("tensor", "hbm")), key=lambda k: kernels[k]["ms_per_step"])
    kd = kernels[dom]
    if kd["bound"] == "tensor":
        roof = {"kernel": dom, "bound": "tensor", "achieved": kd["achieved"], "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                "frac": kd["frac"], "traffic": None, "peak_source": peaks["source"] + " (bf16 sustained)"}
        if args.gemm_mode == 1:
            roof["ceiling_frac"] = 1.0 / 6.0
            roof["ceiling_note"] = ("fp32-faithful 3xTF32: three tf32 products per algorithmic product at half the bf16 rate -> "
                                    "at most 1/6 of the bf16 peak")
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": kd.get("achieved_algorithmic"), "peak": peaks["hbm"], "unit": "GB/s",
                "frac": (kd.get("achieved_algorithmic") or 0.0) / peaks["hbm"], "traffic": None, "peak_source": peaks["source"]}
    roof.update(share_of_step=kd["share_of_step"], launches_per_step=kd["launches_per_step"],
                avg_launch_us=kd["ms_per_step"] * 1e3 / max(kd["launches_per_step"], 1))
    if dom in traffic:
        roof["traffic"] = traffic[dom].get("dram_bytes")
        roof["traffic_note"] = "ncu capture of the largest launch of this group (%s)" % traffic[dom].get("kernel")
    breakdown = {k: round(v["ms"] / nprof, 4) for k, v in prof.items() if v["ms"] > 0}

    # one-by-one mode (B = 1): the reference's own update semantics, one `model.train(uidx)` call per user
    obo = None
    if world == 1 and not args.no_obo:
        n_obo = 256
        for u in range(4):                  # first sights of the shape: plain call, then graph capture
            model.train(np.array([u], dtype=np.int32))
        torch.cuda.synchronize()
        l0, r0 = eng.launch_count(), eng.graph_replays()
        t0 = time.perf_counter()
        for u in range(4, 4 + n_obo):
            model.train(np.array([u], dtype=np.int32))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        obo = {"value": checkins_of(lens_loc[4:4 + n_obo]) / dt, "unit": UNIT, "ms_per_user_call": dt / n_obo * 1e3,
               "kernels_per_call": (eng.launch_count() - l0) / n_obo, "graph_replays": eng.graph_replays() - r0,
               "note": "B=1 calls through the Python class (wall clock incl. host overhead), reference semantics; SIMT "
                       "small-batch recurrence (exact fp32) + CUDA-graph replay"}

    micro = hbm_microbench(eng, dev, peaks) if (world == 1 and not args.no_micro) else None
    parity = parity_block(ds, st, cfg, local_rank) if (world == 1 and not args.no_parity) else None

    cpu = cpu_o = None
    if world == 1 and not args.no_cpu_baseline:
        cb = min(args.cpu_batch, U)
        v, sec = cpu_steps(ds, st, cb, 3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "3 steps x %d users of the %d-user step (numpy oracle, float32, BLAS threads = all cores)" % (cb, B)}
        cpu_o = cpu_obo(ds, st, 256, budget_s=args.cpu_obo_s)

    h2d = 4 * B * seq * 4 + B * 4
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32" if args.gemm_mode in (0, 1) else "tf32",
        "data": "synthetic",
        "config": make_config(args.config, cfg, args.batch, world, args.scaling),
        "engine": {"gemm": {0: "fp32 FMA", 1: "tcgen05 3xTF32 (fp32-faithful), fp32 accumulate in TMEM", 2: "tcgen05 1xTF32"}[args.gemm_mode],
                   "gemm_mode": args.gemm_mode, "fused_recurrence": bool(args.fused) and args.gemm_mode != 0,
                   "fused_cluster": args.fused_cluster or "auto", "check_ins_per_step": done_all / K,
                   "parallelism": "1 GPU" if world == 1 else
                   ("dp%d: users sharded, item table row-sharded (row %% %d); rows gathered from the owners' shards and "
                    "row-gradients pulled from the peers' outboxes by kernels over NVLink peer memory; dense gradients "
                    "all-reduced (NCCL)" if args.peer else
                    "dp%d: users sharded, item table row-sharded (row %% %d) with NCCL all-to-all of rows / row-gradients, "
                    "dense gradients all-reduced") % (world, world)},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 40,
                "ms_per_step": ms_e2e_max / K},
        "gpu_launches": launches_all,
        "roofline": roof, "kernels": kernels, "hbm_microbench": micro, "kernel_ms_per_step": breakdown,
        "sustained": sustained, "parity": parity,
        "cpu_baseline": cpu, "cpu_baseline_obo": cpu_o, "obo_mode": obo, "clocks": clocks, "final_loss": float(losses[-1]),
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch users per GPU per step; strong: --batch users per step split over the GPUs")
    ap.add_argument("--batch", type=int, default=4096, help="users per step per GPU (weak) / per step (strong)")
    ap.add_argument("--c5-items", type=int, default=0, help="c5 smoke runs: catalogue size override (0 = the config's 10M)")
    ap.add_argument("--c5-seq", type=int, default=0, help="c5 smoke runs: sequence length override (0 = the config's 256)")
    ap.add_argument("--positions", type=int, default=4, help="c3 / c4: consecutive positions of every user in one mini-batch step")
    ap.add_argument("--cpu-batch", type=int, default=512, help="users per CPU-oracle step")
    ap.add_argument("--gemm-mode", type=int, default=1,
                    help="0 fp32 FMA, 1 tcgen05 3xTF32 (fp32-faithful, default), 2 tcgen05 1xTF32")
    ap.add_argument("--fused", type=int, default=1, help="1 = persistent fused recurrence kernel (tensor-core modes)")
    ap.add_argument("--fused-cluster", type=int, default=0,
                    help="CTAs per 128 users in the fused recurrence kernels: 0 auto (default), 1, 2, 4")
    ap.add_argument("--peer", type=int, default=1,
                    help="multi-GPU exchange: 1 = NVLink peer-memory kernels (default), 0 = NCCL all-to-all")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="seconds of back-to-back steps for the `sustained` figure (0 = skip)")
    ap.add_argument("--cpu-obo-s", type=float, default=15.0, help="CPU budget of the one-by-one reference-semantics baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-obo", action="store_true")
    ap.add_argument("--no-micro", action="store_true", help="skip the stand-alone gather / scatter HBM micro-benchmark")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.config == "c5" and "--batch" not in " ".join(sys.argv):
        args.batch = 8192 if args.scaling == "strong" else 2048
    if args.config == "c4" and "--batch" not in " ".join(sys.argv):
        args.batch = 128
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
