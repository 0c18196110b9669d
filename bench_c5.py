"""bench.py --config c5: Distance2Pre at |POI| = 10M, |U| = 1M, seq = 256, d = H = 512 (BASELINE.json configs[4]; 8 x B200,
item table row-sharded, dense gradients summed across ranks).

`--scaling strong` (the north_star's target): the GLOBAL batch (`--batch`, default 8192 users per step for this config) is
fixed and split over the N GPUs; `--scaling weak`: `--batch` users per GPU.  Everything is generated on the device
(synth.make_dataset_device / init_state_device: the item table alone is 20.5 GB); a user's sequences and a table row's
values do not depend on N, so every N trains on the same data.  At N = 1 the single-GPU class (SpatialGru) runs; at N > 1
`ShardedSpatialGru` (peer-memory step, csrc/mg_step.cuh), each rank holding rows r, r + N, ... of the table.
At H = 512 the recurrence runs as two tcgen05 GEMM launches per time step (the fused persistent kernels need H <= 128)."""
import json
import os
import sys
import time

import numpy as np

import bench as B0


def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import poi_b200  # noqa: F401
    from poi_b200 import synth
    from poi_b200.engine import Engine
    cfg = dict(synth.CONFIGS["c5"])
    if args.c5_items:
        cfg["n_item"] = args.c5_items                   # smaller catalogue for smoke runs (reported in the workload string)
    if args.c5_seq:
        cfg["seq"] = args.c5_seq
    weak = args.scaling == "weak"
    B_loc = args.batch if weak else max(1, args.batch // world)
    Bg = B_loc * world
    W, K = max(args.warmup, 3), max(args.steps, 1)
    n_steps = W + K + 1 + K + 2                          # resident timed + e2e + profiled
    I, d, seq, U = cfg["n_item"], cfg["d"], cfg["seq"], cfg["n_user"]
    if n_steps * Bg > U:
        raise SystemExit("bench c5: %d steps x %d users exceed |U| = %d" % (n_steps, Bg, U))
    dev = torch.device("cuda", local_rank)
    eng = Engine.get(local_rank)
    # step s trains global users [s Bg, (s+1) Bg); rank r takes the r-th slice of each block
    mine = np.concatenate([np.arange(s * Bg + rank * B_loc, s * Bg + (rank + 1) * B_loc) for s in range(n_steps)])
    t0 = time.time()
    ds = synth.make_dataset_device(eng, U, I, seq, users=mine, UD=cfg["UD"], dd=cfg["dd"])
    D = ds["dist_num"]
    rows = torch.arange(rank, I + 1, world, device=dev) if world > 1 else None
    st = synth.init_state_device(I, d, D, dev, rows=rows)
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    U_loc = ds["n_user"]
    if world == 1:
        from poi_b200.public.GRU_Spatial import SpatialGru
        tes = np.full((1, 1), I, dtype=np.int32)
        init = dict(st); init["trained_items"] = torch.zeros((1, d), device=dev); init["trained_users"] = torch.zeros((1, d), device=dev)
        init["trained_dists"] = torch.zeros((1, d), device=dev)
        model = SpatialGru([ds["P"], ds["lens"], ds["Q"]], [tes, np.ones_like(tes), tes], [ds["DP"], np.full_like(tes, D), ds["DQ"]],
                           [B0.ALPHA, B0.LAM], U_loc, I, [D, cfg["dd"] / 1000.0], d, d, init=init, device=local_rank)
    else:
        from poi_b200.dist import ShardedSpatialGru
        model = ShardedSpatialGru([ds["P"], ds["lens"], ds["Q"]], [ds["DP"], ds["DQ"]], [B0.ALPHA, B0.LAM], I, D, d, d, st,
                                  device=local_rank, lt_is_shard=True, peer=bool(args.peer), max_batch=B_loc)
    del st
    eng.set_gemm_mode(args.gemm_mode)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def users_of(s):
        return np.arange(s * B_loc, (s + 1) * B_loc, dtype=np.int32)

    def step_resident(s):
        return model.train(users_of(s))

    pin = {}

    def step_host_rows(s):
        if s not in pin:
            sl = slice(s * B_loc, (s + 1) * B_loc)
            pin[s] = tuple(ds[k][sl].cpu().pin_memory() for k in ("P", "Q", "DP", "DQ", "lens"))
        return model.train_host_rows(*pin[s])

    def timed(fn, n_warm, n, first):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for i in range(n_warm):
            fn(first + i)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = eng.launch_count(); losses = []
        for i in range(n):
            flush.fill_(i & 0xff)
            ev[i][0].record()
            losses.append(fn(first + n_warm + i)[0])
            ev[i][1].record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        return sum(a.elapsed_time(b) for a, b in ev), eng.launch_count() - l0, losses

    sampler = B0.ClockSampler(local_rank); sampler.start()
    ms, launches, losses = timed(step_resident, W, K, 0)
    for s in range(W + K, W + K + 1 + K):
        sl = slice(s * B_loc, (s + 1) * B_loc)
        pin[s] = tuple(ds[k][sl].cpu().pin_memory() for k in ("P", "Q", "DP", "DQ", "lens"))
    ms_e2e, _, _ = timed(step_host_rows, 1, K, W + K)
    clocks = sampler.stop()
    eng.kprof_reset(); eng.kprof_enable(True)
    nprof = 2
    for i in range(nprof):
        step_resident(W + K + 1 + K + i)
    prof = eng.kprof_get(); eng.kprof_enable(False)

    def red(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())
    mx = (lambda x: red(x, dist.ReduceOp.MAX)) if dist else (lambda x: x)
    sm = (lambda x: red(x, dist.ReduceOp.SUM)) if dist else (lambda x: x)
    ms_max, ms_e2e_max, launches_all = mx(ms), mx(ms_e2e), int(sm(launches))
    mem = mx(torch.cuda.max_memory_allocated(dev) / 2**30)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = B0.load_peaks()
    ci_step = Bg * (seq - 1)
    value = ci_step * K / (ms_max * 1e-3)
    kernels = B0.kernel_table(prof, nprof, peaks, args.gemm_mode)
    dom = max((k for k in kernels if kernels[k]["bound"] == "tensor"), key=lambda k: kernels[k]["ms_per_step"])
    kd = kernels[dom]
    roof = {"kernel": dom, "bound": "tensor", "achieved": kd["achieved"], "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": kd["frac"],
            "traffic": None, "peak_source": peaks["source"] + " (bf16 sustained)", "share_of_step": kd["share_of_step"],
            "ceiling_frac": 1.0 / 6.0 if args.gemm_mode == 1 else None,
            "whole_step_tflops_per_gpu": ci_step * 14.77e6 / world / (ms_max / K * 1e-3) / 1e12,
            "note": "algorithmic flops per check-in 14.77 MFLOP (SURVEY.md 8d); 3xTF32 executes three times that on the tensor pipe"}
    wl = B0.make_config("c5", cfg, args.batch, world, args.scaling)
    line = {"metric": B0.METRIC, "value": value, "unit": B0.UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32" if args.gemm_mode in (0, 1) else "tf32",
            "data": "synthetic (generated on the device)", "config": wl,
            "engine": {"gemm_mode": args.gemm_mode, "recurrence": "two tcgen05 GEMM launches per time step (H = 512 > 128)",
                       "check_ins_per_step": ci_step, "users_per_gpu": B_loc, "setup_s": round(t_setup, 1), "peak_device_mem_gib": round(mem, 1),
                       "parallelism": "1 GPU" if world == 1 else "dp%d: users sharded, item table row-sharded (row %% %d), peer-memory step" % (world, world)},
            "e2e": {"value": ci_step * K / (ms_e2e_max * 1e-3), "unit": B0.UNIT, "h2d_bytes_per_step": B_loc * seq * 16 + B_loc * 4,
                    "d2h_bytes_per_step": 40, "ms_per_step": ms_e2e_max / K},
            "gpu_launches": launches_all, "roofline": roof, "kernels": kernels,
            "kernel_ms_per_step": {k: round(v["ms"] / nprof, 4) for k, v in prof.items() if v["ms"] > 0},
            "clocks": clocks, "final_loss": float(losses[-1])}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
