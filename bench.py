#!/usr/bin/env python
"""bench.py — SAGA sample-updates/s on BASELINE config 2 (binomial lasso, synthetic sparse CSR 1M x 100k, 100 nnz/row),
plus the other two BASELINE metrics (lambda-path fit time, cv_sgdnet fits/s) and the many-fits regime.

A "step" is one pass of the hot path over one batch: ONE SAGA EPOCH (n sample-updates with their lagged prox,
gradient-memory and gradient-average updates) at a fixed lambda of the 100-lambda path, through the C ABI of
libsgdnet_b200.so.

  value        whole-job sample-updates/s, design + state resident in HBM (stepping interface), max over ranks
  e2e          BASELINE metric (ii): ONE sgdnet_fit_sparse call with HOST buffers over the WHOLE 100-lambda path of the
               same workload (CSC in, CSR build, upload, 381 epochs, 100 deviance passes, archives out), as updates/s and
               as seconds
  roofline     HBM roofline of the dominant kernel (saga_sparse_wave_kernel): algorithmic bytes per update
               (12*nnz_row + 8 + 4 + 8*K_y + 16*K = 1236 B) x updates per launch / kernel time
  roofline_passes   the streaming kernel of the path (per-lambda deviance pass) against the same peak
  batch        many concurrent fits on one GPU (128 lasso fits of a 200k x 50k design, one sgdnet_fit_batch_sparse call)
  cv           BASELINE metric (iii) / config 5 at FULL spec: 10-fold x 5-alpha cv_sgdnet, binomial sparse 500k x 50k,
               100 lambdas, thresh 1e-3 - all 55 fits dealt to the N ranks, one NCCL all_gather of the score rows
  dense        configs 3 and 4 (dense multinomial 60000 x 784 K=10; dense mgaussian 200000 x 2000 K=4): updates/s of
               saga_dense_cluster_kernel at full size
  cpu_baseline the reference's CPU path (oracle/_ref and the restated oracle, the faster of the two; g++ -O2, 1 thread:
               the reference is single-threaded) on a bounded sample of the same workload, timed on this box

One process per GPU. A single fit does not shard (the solver is a serial recurrence); with --gpus N every rank runs an
independent fit of the same shape for `value` / `e2e` (weak scaling, no data-path collective), while `cv` shards its 55
fits over the ranks. `--impl reference` times the CPU arm instead, on the GPU arm's own workload and size.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # per-fit streams of batches (engine.cu); before CUDA starts

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NNZ_ROW = 100
OUT = sys.stdout
B_UPD = 12 * NNZ_ROW + 8 + 4 + 8 * 1 + 16 * 1          # 1236 B per sample-update
PARTS = ("e2e", "batch", "cv", "passes", "dense", "cpu")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", dest="n", type=int, default=1_000_000, help="n (rows of X); the metric is quoted at the default")
    ap.add_argument("--cols", dest="p", type=int, default=100_000, help="p (columns of X)")
    ap.add_argument("--lambda-ind", type=int, default=30)
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU time of the bounded oracle samples")
    ap.add_argument("--e2e-epochs", type=int, default=0,
                    help="0 (default): the e2e leg is the whole 100-lambda path (381 epochs); > 0: that many epochs at one lambda")
    ap.add_argument("--skip", default="", help="comma list of parts to leave out: " + ",".join(PARTS))
    ap.add_argument("--cv-nlambda", type=int, default=100)
    ap.add_argument("--cv-maxit", type=int, default=1000)
    ap.add_argument("--batch-fits", type=int, default=128)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    a.skip = set(s for s in a.skip.split(",") if s)
    if a.no_cpu:
        a.skip.add("cpu")
    if a.no_e2e:
        a.skip.add("e2e")
    return a


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [t.strip() for t in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(n, p):
    from sgdnet_b200 import synth
    x, y = synth.binomial_sparse(n, p, NNZ_ROW, seed=1002)
    return x, y


def control_for(lib, n_lambda=100, lambda_=None, maxit=1000, standardize=False):
    from sgdnet_b200 import api
    return api.build_control("binomial", 1, alpha=1.0, nlambda=n_lambda, lambda_min_ratio=1e-4, lambda_=lambda_,
                             maxit=maxit, standardize=standardize, intercept=True, thresh=1e-3, standardize_response=False,
                             debug=False)


def cpu_arms():
    """The CPU implementations that can be timed: oracle/_ref (the reference's own src/sgdnet.cpp compiled against the
    Rcpp/Eigen stand-in, kind "reference") when it was built, and the restated oracle in its libm arithmetic (kind
    "port"). The two produce identical bits (tests/test_ref_cpu.py); the faster one is reported as the baseline."""
    from oracle_lib import load_oracle, load_reference_build
    arms = []
    ref = load_reference_build()
    if ref is not None:
        ref.lib.ref_set_force_debug(0)      # per-epoch loss passes are only needed for the tests' epoch counts
        arms.append((ref, "reference", "reference src/sgdnet.cpp + saga-sparse.h compiled unmodified (g++ -O2, no FMA contraction) against "
                     "the Rcpp/Eigen stand-in of oracle/refbuild (no R/Eigen in the image), 1 thread: the reference is single-threaded"))
    oracle = load_oracle()
    oracle.lib.oracle_set_arith(0)      # std::exp/std::log, sequential sums: the most literal reading of the reference
    arms.append((oracle, "port", "CPU oracle restatement (libm arithmetic), g++ -O2, 1 thread: the reference is single-threaded"))
    return arms


def time_cpu_epochs(lib, x, y, lam, first, last, standardize=False):
    """solver seconds of epochs first+1 .. last of a `last`-epoch fit at one lambda (same seed: the shorter fit is a
    prefix of the longer one)."""
    def run(epochs):
        ctl, keep = control_for(lib, 1, [lam], maxit=epochs, standardize=standardize)
        ctl.tol = 0.0                                   # never converge early: exactly `epochs` epochs
        t0 = time.time()
        raw = lib.fit(x, y.reshape(-1, 1), ctl, lib.rng_from_seed(1))
        return raw, time.time() - t0
    raw, wall = run(last)
    raw_w = run(first)[0] if first > 0 else None
    return raw.seconds_solver - (raw_w.seconds_solver if raw_w is not None else 0.0), raw, wall


def run_reference(args, rank, world):
    """CPU arm on the GPU arm's own workload (same n, p, lambda): `steps` epochs after `warmup` epochs, one thread."""
    if rank != 0:
        return
    n = args.n
    # which of the two CPU implementations is faster here: a short probe on a small sample of the same generator
    xs, ys = make_workload(min(n, 100_000), args.p)
    lam_s = path_lambda(xs, ys, args.lambda_ind)
    tried = []
    for arm in cpu_arms():
        s_probe, _, _ = time_cpu_epochs(arm[0], xs, ys, lam_s, 1, 2)
        tried.append({"kind": arm[1], "probe_updates_per_s": xs.shape[0] / s_probe, "arm": arm})
    lib, kind, note = max(tried, key=lambda d: d["probe_updates_per_s"])["arm"]
    for d in tried:
        del d["arm"]
    # ... and that one on the GPU arm's own workload and size
    x, y = make_workload(n, args.p)
    lam = path_lambda(x, y, args.lambda_ind)
    solver_steps, raw, wall = time_cpu_epochs(lib, x, y, lam, args.warmup, args.warmup + args.steps)
    ups = n * args.steps / solver_steps
    line = {
        "impl": "reference", "metric": "SAGA sample-updates/s", "value": ups, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": solver_steps / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args),
        "cpu_baseline": {"value": ups, "unit": "updates/s", "cores": 1, "kind": kind,
                         "sample": f"last {args.steps} of {raw.npasses} epochs of the full workload (n={n}, p={args.p}, {NNZ_ROW} nnz/row) at "
                                   f"lambda[{args.lambda_ind}]; " + note, "arms_timed": tried},
        "e2e": {"value": n * raw.npasses / wall, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


def path_lambda(x, y, ind):
    """lambda[ind] of the automatic 100-lambda path of this workload (families.h:203-220 with standardize = FALSE; the rest
    of the path is LogSpace, math.h:42-56)."""
    xc = x.tocsc()
    yb = (y - y.mean()) / y.std()
    lmax = y.std() * np.abs(xc.T @ yb).max() / x.shape[0]
    return float(np.exp(np.log(lmax) + ind * (np.log(lmax * 1e-4) - np.log(lmax)) / 99.0))


def config_dict(args):
    return {"workload": "BASELINE config 2: binomial lasso (alpha=1), sparse CSR 1M x 100k, 100 nnz/row, standardize=FALSE, "
                        "intercept=TRUE; step = one SAGA epoch at lambda[%d] of the 100-lambda path" % args.lambda_ind,
            "n": int(args.n), "p": args.p, "nnz_row": NNZ_ROW, "family": "binomial", "penalty": "lasso", "fits_per_gpu": 1,
            "l2_policy": "inputs larger than L2: 1.2 GB CSR per pass vs 126 MB L2, random row order"}


# ------------------------------------------------------------------------------------------------- parts of the b200 arm
def part_batch(lib, args):
    """Many independent fits on one GPU (alpha grids / cv folds / bootstrap replicates): F lasso fits of one sparse design,
    different seeds, through ONE sgdnet_fit_batch_sparse call with host buffers; every fit is its own pipeline."""
    from sgdnet_b200 import _abi, api, synth
    F, n, p, E = args.batch_fits, 200_000, 50_000, 6
    x, y = synth.binomial_sparse(n, p, NNZ_ROW, seed=1002)
    yc = y - y.mean()
    lam = [float(np.abs(x.T @ yc).max() / n) * 0.05]
    m = _abi.CscMatrix.from_any(x)

    def run(f):
        specs, keeps = [], []
        for k in range(f):
            ctl, keep = api.build_control("binomial", 1, alpha=1.0, nlambda=1, lambda_min_ratio=1e-4, lambda_=lam, maxit=E,
                                          standardize=False, intercept=True, thresh=0.0, standardize_response=False, debug=False)
            keeps.append(keep)
            specs.append(dict(train_rows=None, test_rows=None, control=ctl, rng=lib.rng_from_seed(100 + k)))
        t0 = time.perf_counter()
        raws, _ = lib.fit_batch(m, y.reshape(-1, 1), specs)
        return raws, time.perf_counter() - t0
    run(2)                                              # CUDA module load and allocator warm-up are not the batch's
    raws, wall = run(F)
    updates = sum(int(r.npasses) for r in raws) * n
    solver = max(r.seconds_solver for r in raws)
    return {"workload": f"{F} concurrent lasso fits of binomial sparse {n}x{p}, {NNZ_ROW} nnz/row, {E} epochs each, one "
                        "sgdnet_fit_batch_sparse call (host CSC in, results out)",
            "fits": F, "updates": updates, "wall_s": wall, "solver_s_longest_fit": solver, "setup_s": max(r.seconds_setup for r in raws),
            "wall_over_solver": wall / solver, "agg_updates_per_s_wall": updates / wall, "agg_updates_per_s_solver": updates / solver,
            "algorithmic_GBps_wall": updates / wall * B_UPD / 1e9}


def part_cv(lib, args, shard, dist, world):
    """BASELINE config 5 at full spec. Returns the `cv` object of the JSON line after one all_gather of the score rows
    (inside cv_sgdnet) and two all_reduces of the counters reported here."""
    import torch
    import sgdnet_b200 as sg
    from sgdnet_b200 import synth
    n, p = 500_000, 50_000
    x, y = synth.binomial_sparse(n, p, 50, seed=1005)
    foldid = (np.random.Generator(np.random.PCG64(1005)).permutation(n) % 10) + 1
    alphas = [0.0, 0.25, 0.5, 0.75, 1.0]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cv = sg.cv_sgdnet(x, y, family="binomial", alpha=alphas, foldid=foldid, nlambda=args.cv_nlambda, standardize=False,
                      maxit=args.cv_maxit, thresh=1e-3, seed=1000, backend=lib, shard=shard)
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    full = [cv.fits[i] for i in cv.owned_full]
    folds = [f for f in cv.fold_fits if f is not None]
    mine = [(int(f.npasses) * n, f.raw.seconds_solver) for f in full] + [(int(f.npasses) * f.nobs, f.raw.seconds_solver) for f in folds]
    upd = sum(u for u, _ in mine)
    t = torch.tensor([float(upd), float(len(mine))], dtype=torch.float64, device="cuda")
    tm = torch.tensor([wall, max([s for _, s in mine] or [0.0]), float(max([u for u, _ in mine] or [0]))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    upd, nfits, wall, longest, longest_upd = float(t[0]), int(t[1]), float(tm[0]), float(tm[1]), float(tm[2])
    limiter = ("the longest single fit: a full-data fit is one serial SAGA recurrence on one SM; its %.1f s of solver time are "
               "%.0f %% of the wall" % (longest, 100 * longest / wall))
    return {"workload": "BASELINE config 5: cv_sgdnet, 10 folds x 5 alphas {0,.25,.5,.75,1}, binomial sparse 500k x 50k, 50 nnz/row, "
                        f"nlambda={args.cv_nlambda}, thresh=1e-3, maxit={args.cv_maxit}, standardize=FALSE, train-on-one-fold (R/cv_sgdnet.R:182-183), "
                        "host buffers in, score rows gathered with one all_gather",
            "n_gpus": world, "fits": nfits, "wall_s": wall, "fits_per_s": nfits / wall, "updates": upd, "agg_updates_per_s": upd / wall,
            "longest_fit_solver_s": longest, "longest_fit_updates": longest_upd, "limiter": limiter,
            "alpha_min": float(cv.alpha_min), "lambda_min": float(cv.lambda_min),
            "sharding": "all 55 fits dealt longest-first over the ranks; per-fit seeds 1000 + fit index"}


def part_dense(lib):
    """Configs 3 and 4 at full size: epochs of the dense cluster kernel (saga_dense_cluster.cu; p >= 512) through the stepping
    interface (design resident)."""
    from sgdnet_b200 import _abi, api, synth
    out = {}
    for name, gen, fam, K, alpha, epochs in (("config3_multinomial_60000x784_K10", lambda: synth.multinomial_dense(60_000, 784, 10, seed=1003), "multinomial", 10, 0.8, 3),
                                             ("config4_mgaussian_200000x2000_K4", lambda: synth.mgaussian_dense(200_000, 2000, 4, seed=1004), "mgaussian", 4, 1.0, 2)):
        x, y = gen()
        n, p = x.shape
        ymat = np.asarray(y, dtype=np.float64).reshape(n, -1)
        ctl, keep = api.build_control(fam, K, alpha=alpha, nlambda=100, lambda_min_ratio=1e-4, lambda_=None, maxit=1000, standardize=True,
                                      intercept=True, thresh=1e-3, standardize_response=False, debug=False)
        xa = np.asfortranarray(x, dtype=np.float64)
        ya = np.asfortranarray(ymat)
        sess = C.c_void_p()
        lib.check(lib.sym("session_create_dense")(_abi._ptr(xa, _abi.c_double_p), C.c_int64(n), C.c_int64(p), _abi._ptr(ya, _abi.c_double_p),
                                                  C.c_int32(ya.shape[1]), C.byref(ctl), C.byref(sess)), "session_create_dense")
        rng = lib.rng_from_seed(1)
        ms = C.c_float(0)
        times = []
        for _ in range(epochs):
            lib.check(lib.sym("session_run_epochs")(sess, 30, 1, C.byref(rng), C.byref(ms)), "run_epochs")
            times.append(ms.value)
        lib.sym("session_destroy")(sess)
        t = min(times[1:]) * 1e-3
        b_upd = 8 * p + 4 + 8 * (K if fam == "mgaussian" else 1) + 16 * K
        out[name] = {"updates_per_s": n / t, "epoch_ms": t * 1e3, "bytes_per_update": b_upd, "algorithmic_GBps": n / t * b_upd / 1e9,
                     "kernel": "saga_dense_cluster_kernel (thread-block cluster, one per fit)",
                     "cycles_per_update_at_1965MHz": t / n * 1.965e9}
        del x, xa
    # sparse + standardize = TRUE (virtual centring: the reference's O(p) sweeps per update, saga_sparse_centred.cu) at
    # reduced p, against the oracle's portable restatement on one host core
    x, y = synth.binomial_sparse(20_000, 2000, 16, seed=1012)
    m = _abi.CscMatrix.from_any(x)
    n, p = m.shape
    ya = np.ascontiguousarray(y.reshape(-1, 1))
    ctl, keep = api.build_control("binomial", 1, alpha=1.0, nlambda=100, lambda_min_ratio=1e-4, lambda_=None, maxit=1000, standardize=True,
                                  intercept=True, thresh=1e-3, standardize_response=False, debug=False)
    sess = C.c_void_p()
    lib.check(lib.sym("session_create_sparse")(_abi._ptr(m.i, _abi.c_int32_p), _abi._ptr(m.p, _abi.c_int32_p), _abi._ptr(m.x, _abi.c_double_p),
                                               C.c_int64(n), C.c_int64(p), _abi._ptr(ya, _abi.c_double_p), C.c_int32(1), C.byref(ctl),
                                               C.byref(sess)), "session_create_sparse")
    rng = lib.rng_from_seed(1)
    ms = C.c_float(0)
    times = []
    for _ in range(3):
        lib.check(lib.sym("session_run_epochs")(sess, 30, 1, C.byref(rng), C.byref(ms)), "run_epochs")
        times.append(ms.value)
    lib.sym("session_destroy")(sess)
    t = min(times[1:]) * 1e-3
    out["config2_standardized_sparse_20000x2000"] = {"updates_per_s": n / t, "epoch_ms": t * 1e3, "kernel": "saga_sparse_centred_kernel (one CTA, owner "
                                                     "computes; virtual centring touches all p coefficients on every update, as in the reference)",
                                                     "cycles_per_update_at_1965MHz": t / n * 1.965e9}
    try:
        from oracle_lib import load_oracle
        lam = path_lambda(x, y, 30)
        cpu_s, _, _ = time_cpu_epochs(load_oracle(), x, y, lam, 1, 3, standardize=True)
        out["config2_standardized_sparse_20000x2000"]["cpu_oracle_updates_per_s_one_core"] = n * 2 / cpu_s
    except Exception as exc:      # noqa: BLE001
        out["config2_standardized_sparse_20000x2000"]["cpu_oracle_error"] = str(exc)
    return out


def part_passes(lib, sess, n, nnz, peak, rng):
    """The per-lambda deviance pass (rescale + loss pass + finish_lambda kernels) on config 2's design: CUDA-event time of
    sgdnet_session_finish_lambda; algorithmic bytes 12*nnz + 16*n + 8*n. Two states of the coefficients:
    as a lasso path sees them (few nonzero weights: the nonzero bitmap filters the gathers and the pass streams X at HBM
    speed) and all weights live (one 8-byte gather per nonzero of X: bound by the LSU's one gather wavefront per cycle
    per SM, 148 x 1.965 GHz x 12 B = 3.5 TB/s of algorithmic bytes at most)."""
    ms = C.c_float(0)
    bytes_pass = 12 * nnz + 16 * n + 8 * n

    def timed(lambda_ind):
        times = []
        for _ in range(6):
            lib.check(lib.sym("session_finish_lambda")(sess, lambda_ind, C.byref(ms)), "finish_lambda")
            times.append(ms.value)
        return float(np.median(times[1:])) * 1e-3

    def entry(t, state):
        return {"kernel": "rescale_kernel + loss_pass_tiles_kernel + finish_lambda_kernel", "bound": "hbm", "achieved": bytes_pass / t / 1e9,
                "peak": peak, "unit": "GB/s", "frac": bytes_pass / t / 1e9 / peak, "launch_ms": t * 1e3, "bytes_per_pass": bytes_pass,
                "workload": "config 2 design (1M x 100k, 1e8 nonzeros), one lambda; " + state}
    out = {"deviance_pass_all_weights_live": entry(timed(30), "coefficients after the timed epochs at lambda[30] from a cold start "
                                                              "(nearly every weight nonzero: gather-bound)")}
    # a state as the warm-started path sees it: a few epochs at a strong penalty zero most weights
    lib.check(lib.sym("session_run_epochs")(sess, 3, 3, C.byref(rng), C.byref(ms)), "run_epochs")
    out["deviance_pass"] = entry(timed(3), "coefficients after 3 further epochs at lambda[3] (lasso-sparse weights, as along the path)")
    return out


def part_cpu_cv(args):
    """CPU legs of config 5: per-core rate of the reference's CPU path on a fold-shaped fit,
    alone and with min(55, nproc) concurrent worker processes."""
    import concurrent.futures as cf
    nproc = os.cpu_count() or 1
    workers = min(55, nproc)
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from sgdnet_b200 import synth, api\n"
            "from oracle_lib import load_oracle, load_reference_build\n"
            "lib = load_reference_build() or load_oracle()\n"
            "getattr(lib.lib, 'ref_set_force_debug', lambda v: None)(0)\n"
            "getattr(lib.lib, 'oracle_set_arith', lambda v: None)(0)\n"
            "x, y = synth.binomial_sparse(50_000, 50_000, 50, seed=1005)\n"
            "ctl, keep = api.build_control('binomial', 1, alpha=0.5, nlambda=1, lambda_min_ratio=1e-4, lambda_=[1e-4], maxit=%d, standardize=False, "
            "intercept=True, thresh=0.0, standardize_response=False, debug=False)\n"
            "raw = lib.fit(x, y.reshape(-1, 1), ctl, lib.rng_from_seed(1))\n"
            "print(json.dumps({'updates': 50_000 * int(raw.npasses), 'solver_s': raw.seconds_solver}))\n")
    epochs = max(2, int(args.cpu_seconds * 0.35e6 / 50_000))

    def one(_):
        out = subprocess.run([sys.executable, "-c", code % (ROOT, os.path.join(ROOT, "tests"), epochs)], capture_output=True, text=True)
        return json.loads(out.stdout.strip().splitlines()[-1])
    alone = one(0)
    with cf.ThreadPoolExecutor(workers) as ex:
        together = list(ex.map(one, range(workers)))
    rate1 = alone["updates"] / alone["solver_s"]
    rate_p = float(np.mean([r["updates"] / r["solver_s"] for r in together]))
    return {"cores_available": nproc, "workers": workers, "updates_per_s_one_core_alone": rate1,
            "updates_per_s_per_core_with_all_workers_busy": rate_p,
            "sample": f"{epochs} epochs of one fold-shaped fit (50k x 50k, 50 nnz/row, elastic net) per process; oracle/_ref when built, else the restated oracle"}


def main():
    args = parse()
    # ONE JSON line on stdout: libraries (NCCL prints its version banner to fd 1) are sent to stderr
    global OUT
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    import sgdnet_b200 as sg
    from sgdnet_b200 import _abi
    from sgdnet_b200.shard import Shard
    lib = sg.product()
    lib.check(lib.sym("set_device")(local), "set_device")
    peak, peak_src = peaks()

    x, y = make_workload(args.n, args.p)
    m = _abi.CscMatrix.from_any(x)
    n, p = m.shape
    ya = np.ascontiguousarray(y.reshape(-1, 1), dtype=np.float64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: resident stepping interface
    ctl, keep = control_for(lib)
    sess = C.c_void_p()
    xargs = [_abi._ptr(m.i, _abi.c_int32_p), _abi._ptr(m.p, _abi.c_int32_p), _abi._ptr(m.x, _abi.c_double_p), C.c_int64(n), C.c_int64(p)]
    lib.check(lib.sym("session_create_sparse")(*xargs, _abi._ptr(ya, _abi.c_double_p), C.c_int32(1), C.byref(ctl), C.byref(sess)),
              "session_create_sparse")
    rng = lib.rng_from_seed(1 + rank)
    ms = C.c_float(0)
    step = lambda: lib.check(lib.sym("session_run_epochs")(sess, args.lambda_ind, 1, C.byref(rng), C.byref(ms)), "run_epochs")
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    kernel_ms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        kernel_ms.append(ms.value)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    # device time of the timed region = sum of the CUDA-event brackets around the epoch kernels (the fit's own stream);
    # wall (host clock around barrier + sync) also holds the host visit between launches. The next epoch's sampling
    # indices and conflict codes are produced on the fit's second stream WHILE the solver runs: in the wall, not in the
    # brackets.
    dev_s = sum(kernel_ms) * 1e-3
    t = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall = float(t[0]), float(t[1])
    updates = n * args.steps * world
    value = updates / wall
    # per epoch: solver + (second stream) MT index generation + conflict codes + lag-scaling table (a no-op for lasso)
    launches = 4 * args.steps

    # ---------------- roofline of the dominant kernel (one launch = one epoch = n updates)
    per_launch_s = (sum(kernel_ms) / len(kernel_ms)) * 1e-3
    achieved = n * B_UPD / per_launch_s / 1e9
    traffic, traffic_src = None, None
    for prof_name in ("r2_wave_ncu_summary.json", "r1_12_wave_fastpath_ncu_summary.json"):
        try:      # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on this workload
            with open(os.path.join(ROOT, "profiles", prof_name)) as fh:
                prof = json.load(fh)
            if n == prof["updates_per_launch"]:
                traffic, traffic_src = prof["dram_bytes_per_launch"], "profiles/" + prof_name
                break
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "saga_sparse_wave_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "bytes_per_update": B_UPD, "updates_per_launch": n, "launch_ms": per_launch_s * 1e3,
                "note": "serial recurrence: one CTA per fit, bounded by the intercept/gradient chain (about 590 cycles per update), not by HBM "
                        "(DESIGN.md); the HBM-bound kernel of the path is in roofline_passes, the many-fits regime in batch / cv"}

    roofline_passes = None
    if "passes" not in args.skip and world == 1:
        try:
            roofline_passes = part_passes(lib, sess, n, int(m.p[-1]), peak, rng)
        except Exception as exc:      # noqa: BLE001
            roofline_passes = {"error": f"{type(exc).__name__}: {exc}"}
    lib.sym("session_destroy")(sess)

    # ---------------- e2e: sgdnet_fit_sparse with host buffers
    e2e = None
    if "e2e" not in args.skip:
        if args.e2e_epochs > 0:
            lam = path_lambda(x, y, args.lambda_ind)
            ctl2, keep2 = control_for(lib, 1, [lam], maxit=max(2, args.e2e_epochs))
            ctl2.tol = 0.0
            what = "%d epochs at one lambda" % max(2, args.e2e_epochs)
        else:
            ctl2, keep2 = control_for(lib)
            what = "the whole 100-lambda path (thresh 1e-3, warm starts, 100 deviance passes)"
        barrier()
        t0 = time.perf_counter()
        raw = lib.fit(m, ya, ctl2, lib.rng_from_seed(1 + rank))
        barrier()
        e2e_wall = time.perf_counter() - t0
        tt = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_wall = float(tt[0])
        h2d = m.i.nbytes + m.x.nbytes + 16 * n + ya.nbytes + 2512      # design, response, generator state (indices are drawn on the device)
        d2h = raw.beta.nbytes + raw.a0.nbytes + raw.dev_ratio.nbytes + 8 * len(raw.lambda_)
        e2e = {"value": n * raw.npasses * world / e2e_wall, "unit": "updates/s", "h2d_bytes_per_step": int(h2d / raw.npasses),
               "d2h_bytes_per_step": int(d2h / raw.npasses), "epochs": int(raw.npasses), "wall_s": e2e_wall,
               "lambda_path_fit_seconds": e2e_wall if args.e2e_epochs == 0 else None,
               "setup_s": raw.seconds_setup, "solver_s": raw.seconds_solver, "deviance_s": raw.seconds_deviance,
               "note": "one sgdnet_fit_sparse call per rank: host CSC in, CSR build + upload, " + what + ", archives out"}
    del x, m

    # ---------------- many fits on one GPU; dense configs (N = 1 only)
    def guarded(part, *a):      # a failing side measurement must not cost the run its headline line
        try:
            return part(*a)
        except Exception as exc:      # noqa: BLE001
            return {"error": f"{type(exc).__name__}: {exc}"}
    batch = guarded(part_batch, lib, args) if ("batch" not in args.skip and world == 1) else None
    dense = guarded(part_dense, lib) if ("dense" not in args.skip and world == 1) else None

    # ---------------- config 5 (cv): sharded over the run's ranks
    cv = None
    if "cv" not in args.skip:
        cv = part_cv(lib, args, Shard.from_torch() if world > 1 else None, dist, world) if world > 1 else guarded(
            part_cv, lib, args, None, dist, world)

    # ---------------- cpu baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and "cpu" not in args.skip:
        n_cpu = int(min(n, max(20_000, 0.4e6 * args.cpu_seconds / 3)))
        xs, ys = make_workload(n_cpu, args.p)
        lam = path_lambda(xs, ys, args.lambda_ind)
        tried, best = [], None
        for lib_c, kind, note in cpu_arms():
            cpu_s, _, _ = time_cpu_epochs(lib_c, xs, ys, lam, 1, 3)      # epochs 2-3 are timed (epoch 1 runs on a cold cache)
            ups = n_cpu * 2 / cpu_s
            tried.append({"kind": kind, "value": ups})
            if best is None or ups > best[0]:
                best = (ups, kind, note)
        cpu = {"value": best[0], "unit": "updates/s", "cores": 1, "kind": best[1],
               "sample": f"epochs 2-3 of a 3-epoch fit of n={n_cpu} rows (same generator, p={args.p}, {NNZ_ROW} nnz/row) at lambda[{args.lambda_ind}]; "
                         + best[2], "arms_timed": tried, "host_cores_available": os.cpu_count()}
        if cv is not None and "error" not in cv:
            legs = part_cpu_cv(args)
            rate_p = legs["updates_per_s_per_core_with_all_workers_busy"]
            legs["sequential_s_extrapolated"] = cv["updates"] / legs["updates_per_s_one_core_alone"]
            legs["parallel_s_extrapolated"] = max(cv["updates"] / (rate_p * legs["workers"]), cv["longest_fit_updates"] / rate_p)
            legs["note"] = ("extrapolated, NOT run: measured per-core rates x the sample-updates of the 55 fits as the B200 run counted them (path "
                            "lengths are equal on both sides by construction, see the parity tests); the parallel leg cannot be shorter than "
                            "its longest fit on one core")
            cv["cpu_legs"] = legs

    if rank == 0:
        line = {"metric": "SAGA sample-updates/s", "value": value, "unit": "updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args),
                "device_ms_per_step": dev_s / args.steps * 1e3, "value_device_only": updates / dev_s,
                "roofline": roofline, "roofline_passes": roofline_passes, "cpu_baseline": cpu, "e2e": e2e, "batch": batch, "cv": cv,
                "dense": dense, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
