"""Timeline of the sparse wavefront kernel (library built with -DSGD_WAVE_TRACE as libsgdnet_b200_trace.so): eight
clock64 events per row for 4096 consecutive rows of an epoch, reduced to where the chain warp's time goes.
Usage: python scripts/wave_trace.py [n] [p]   (writes gpurun_out/wave_trace.npy)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgdnet_b200 import _abi, api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
lib = _abi.Library(os.path.join(ROOT, "sgdnet_b200", "libsgdnet_b200_trace.so"), "sgdnet_")
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
for it in range(3):
    lib.check(lib.sym("session_run_epochs")(sess, 30, 1, C.byref(rng), C.byref(ms)), "run")
R = 4096
buf = (C.c_longlong * (R * 16))()
lib.lib.sgdnet_debug_wave_trace(buf)
a = np.array(buf[:], dtype=np.int64).reshape(R, 16)
np.save(os.path.join(ROOT, "gpurun_out", "wave_trace.npy"), a)
print(f"epoch {ms.value:.1f} ms = {ms.value * 1e-3 * 1.965e9 / n:.0f} cycles/row (trace build), n={n} p={p}")
a = a[8:-8]
C0, C1, C2, W3, W4, W5, W6, W7, need, nr = [a[:, i].astype(np.float64) for i in range(10)]
need = a[:, 8]
period = np.diff(C0)
print(f"chain period (row start to row start): mean {period.mean():.0f}, median {np.median(period):.0f}")
print(f"chain: start->gok published {np.mean(C1 - C0):.0f}; gok published->next operands in hand {np.mean(C2 - C1):.0f}; in hand->next row start {np.mean(C0[1:] - C2[:-1]):.0f}")
# per row t: when did rdy(t) get published relative to the moment the chain could have started it (C1[t-1] + tail)?
lead = C0[1:] - W5[1:]            # > 0: rdy was there before the chain took the row up
print(f"rdy published before the chain takes the row up by (mean) {lead.mean():.0f}; rows where rdy came less than 200 cycles before: {100 * np.mean(lead < 200):.1f} %")
dmin = np.array([(int(v) & 0xffff & -(int(v) & 0xffff)).bit_length() - 1 if (int(v) & 0xffff) else 0 for v in need])
for d in range(0, 8):
    sel = (dmin[1:] == d)
    if sel.sum() == 0:
        continue
    print(f"  nearest needed row d={d}: {100 * sel.mean():5.1f} % of rows, chain period {period[sel].mean():6.0f}, "
          f"rdy->taken up {lead[sel].mean():7.0f}, worker: full->gathered {np.mean((W4 - W3)[1:][sel]):5.0f}, gathered->rdy {np.mean((W5 - W4)[1:][sel]):5.0f}, "
          f"rdy->gok seen {np.mean((W6 - W5)[1:][sel]):5.0f}, gok seen->done {np.mean((W7 - W6)[1:][sel]):5.0f}")
print(f"worker row: full->done mean {np.mean(W7 - W3):.0f}; gok published->seen by the worker {np.mean(W6 - C1):.0f}; rdy published->chain has operands (only rows the chain waited for) "
      f"{np.mean((C2[:-1] - W5[1:])[lead < 50]):.0f}")
S = 8
nxt = W3[S:] - W7[:-S]
print(f"worker: done(t) -> full(t+S) seen {nxt.mean():.0f}")

S10, S11, S12, S13, W14 = [a[:, i].astype(np.float64) for i in (10, 11, 12, 13, 14)]
print(f"scout: period {np.diff(S13).mean():.0f}; full seen->pair barrier {np.mean(S11 - S10):.0f}; barrier->tests done {np.mean(S12 - S11):.0f}; "
      f"tests done->published {np.mean(S13 - S12):.0f}; published(t-1)->full seen(t) {np.mean(S10[1:] - S13[:-1]):.0f}")
print(f"worker: full seen->coded seen {np.mean(W3 - W14):.0f}; scout published -> worker takes the row up {np.mean(W14 - S13):.0f} (negative: the worker waits for the scout)")
