"""Per-phase cycle counters of the sparse wavefront kernel (library built with -DSGD_WAVE_PROF as
sgdnet_b200/libsgdnet_b200_prof.so). Usage: python scripts/wave_prof.py [n] [p]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgdnet_b200 import _abi, api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
lib = _abi.Library(os.path.join(ROOT, "sgdnet_b200", "libsgdnet_b200_prof.so"), "sgdnet_")
x, y = synth.binomial_sparse(n, p, 100, seed=1002)
m = _abi.CscMatrix.from_any(x)
ya = np.ascontiguousarray(y.reshape(-1, 1))
ctl, keep = api.build_control("binomial", 1, alpha=1.0, nlambda=100, lambda_min_ratio=1e-4, lambda_=None, maxit=1000,
                              standardize=False, intercept=True, thresh=1e-3, standardize_response=False, debug=False)
sess = C.c_void_p()
lib.check(lib.sym("session_create_sparse")(_abi._ptr(m.i, _abi.c_int32_p), _abi._ptr(m.p, _abi.c_int32_p),
                                          _abi._ptr(m.x, _abi.c_double_p), C.c_int64(n), C.c_int64(p),
                                          _abi._ptr(ya, _abi.c_double_p), C.c_int32(1), C.byref(ctl), C.byref(sess)), "create")
rng = lib.rng_from_seed(1)
ms = C.c_float(0)
st = (C.c_longlong * 32)()
for it in range(3):
    lib.lib.sgdnet_debug_wave_stall(st, 1)
    lib.check(lib.sym("session_run_epochs")(sess, 30, 1, C.byref(rng), C.byref(ms)), "run")
out = (C.c_longlong * 160)()
lib.lib.sgdnet_debug_wave_prof(out)
a = np.array(out[:], dtype=np.int64).reshape(20, 8)
S = int(os.environ.get("SGDNET_WAVE_WARPS", "8"))
print(f"epoch {ms.value:.1f} ms, {ms.value * 1e-3 * 1.965e9 / n:.0f} cycles/row, S={S}")
print("worker rows each:", n // S)
names = ["wait_full", "issue early+gm", "late waits", "catchup+dot+publish", "wait gok", "row total(after full)", "done chain", "-"]
for w in range(S):
    print(f"worker {w}: " + "  ".join(f"{names[i]}={a[w, i] / (n / S):.0f}" for i in range(7)))
print(f"chain: wait_rdy={a[S + 1, 0] / n:.0f} compute={a[S + 1, 1] / n:.0f} per row")
lib.lib.sgdnet_debug_wave_stall(st, 0)
sa = np.array(st[:], dtype=np.int64).reshape(16, 2)
print("chain stall by nearest-conflict distance of the awaited row (0 = no conflict in window): rows, share of rows, mean wait, share of total wait")
tw = sa[:, 1].sum()
for d in range(16):
    if sa[d, 0]:
        print(f"  d={d:2d}: rows={sa[d,0]:7d} ({100*sa[d,0]/sa[:,0].sum():5.1f}%)  mean wait={sa[d,1]/sa[d,0]:7.0f}  share={100*sa[d,1]/tw:5.1f}%")
pa = (C.c_longlong * 12)()
lib.lib.sgdnet_debug_wave_path(pa)
pa = np.array(pa[:], dtype=np.float64)
if pa[0]:
    print("distance-1 conflict path (mean cycles over %d rows, all epochs): gok->worker wake %.0f | scatter->fdone %.0f | fdone->late wake %.0f | late wake->rdy %.0f | rdy->chain wake %.0f | total %.0f"
          % (pa[0], pa[1] / pa[0], pa[2] / pa[0], pa[3] / pa[0], pa[4] / pa[0], pa[5] / pa[0], pa[6] / pa[0]))
    print("  late wake->rdy split: wake->sum %.0f | butterfly %.0f | dmin+publish %.0f" % (pa[7] / pa[0], pa[8] / pa[0], pa[9] / pa[0]))
