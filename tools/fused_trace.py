"""Timeline of the fused recurrence kernels (gru_fused.cuh): CTA 0 stamps the SM clock at every hand-off
between the MMA thread and the epilogue warps; this script runs c2 steps on an instrumented build
(-DPOI_FUSED_TRACE) and prints the median cycles per phase of a time step.

  python tools/fused_trace.py --build      # here (CPU): compile libpoi_b200_trace.so
  python tools/fused_trace.py              # on the GPU box
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "point-of-interest-recommendation_b200")
TRACE_LIB = os.path.join(PKG, "libpoi_b200_trace.so")

if "--build" in sys.argv:
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-DPOI_FUSED_TRACE", "-o", TRACE_LIB, os.path.join(PKG, "csrc", "poi_engine.cu")]
    print(" ".join(cmd)); subprocess.check_call(cmd); sys.exit(0)

os.environ["POI_B200_LIB"] = TRACE_LIB
import ctypes
import numpy as np
import poi_b200  # noqa
from poi_b200 import synth
from poi_b200._lib import lib
from poi_b200.public.GRU_Spatial import SpatialGru

U, I, seq, d, D, B = 10000, 40000, 32, 128, 200, 4096
ds = synth.make_dataset(U, I, seq)
D = ds["dist_num"]
st = synth.init_state(I, d, d, D)
tes = ds["tes"]
m = SpatialGru([ds["P"], ds["M"], ds["Q"]], [tes, np.ones_like(tes), tes], [ds["DP"], np.full_like(tes, D), ds["DQ"]],
               [0.01, 0.001], U, I, [D, 0.2], d, d, init=st, device=0)
users = np.arange(B, dtype=np.int32)
CL = int(os.environ.get("POI_FUSED_CLUSTER", "0"))
m.engine.set_fused_cluster(CL)
print("#### fused_cluster =", CL or "auto")
for _ in range(3):
    m.train(users)
lib.poi_debug_fused_trace.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
T = seq - 1
for direc in (0, 1):
    lib.poi_debug_fused_trace(direc, None, 0, 1)
m.train(users)
names = {0: ["mma:a_ready#1", "mma:gemm1 issued", "mma:a_ready#2", "mma:gemm2 issued", "epi0:d1_full", "epi0:epi1 done",
             "epi0:d2_full", "epi0:epi2 done", "epi7:d1_full", "epi7:epi1 done", "epi7:d2_full", "epi7:epi2 done",
             "(sum) epi0 wait ax_full", "(sum) mma wait w_full gemm1", "(sum) mma wait w_full gemm2", ""],
         1: ["mma:a_ready da_c", "mma:gemmM issued", "mma:a_ready da_z", "mma:gemmDH1 issued", "mma:a_ready da_r",
             "mma:gemmDH2 issued", "epi0:ddh_full (P start)", "epi0:P done", "epi0:dm_full", "epi0:M1 done", "epi0:M2 done",
             "epi0:dh1_done", "epi0:M3 done", "(sum) epi0 wait in_full", "(sum) mma wait w_full", ""]}
for direc in (0, 1):
    buf = np.zeros(T * 16, dtype=np.int64)
    lib.poi_debug_fused_trace(direc, buf.ctypes.data, T * 16, 0)
    tr = buf.reshape(T, 16)
    print("==== %s recurrence, CTA 0, T=%d: cycles relative to the step's first MMA-side stamp (median over steps 2..T-2)" %
          ("forward" if direc == 0 else "backward", T))
    sel = tr[2:T - 2]
    base = sel[:, 0:1]
    nstamp = 12 if direc == 0 else 13
    order = np.argsort(np.median(sel[:, :nstamp] - base, axis=0))
    for s in order:
        print("  %-28s %8.0f" % (names[direc][s], np.median(sel[:, s] - base[:, 0])))
    for s in range(nstamp, 15):
        print("  %-28s %8.0f" % (names[direc][s], np.median(sel[:, s])))
    step = np.diff(tr[:, 0])
    print("  step period (mma a_ready#1 to next): median %.0f cycles, min %.0f max %.0f" % (np.median(step[1:-1]), step[1:-1].min(), step[1:-1].max()))
