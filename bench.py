#!/usr/bin/env python
"""bench.py — SAGA sample-updates/s on BASELINE config 2 (binomial lasso, synthetic sparse CSR 1M x 100k, 100 nnz/row).

A "step" is one pass of the hot path over one batch: ONE SAGA EPOCH (n sample-updates with their lagged prox,
gradient-memory and gradient-average updates) at a fixed lambda of the 100-lambda path, through the C ABI of
libsgdnet_b200.so.

  value        whole-job sample-updates/s, design + state resident in HBM (stepping interface), CUDA-event timed,
               max over ranks
  e2e          the same metric through the reference-facing call sgdnet_fit_sparse with HOST buffers: CSC -> device,
               setup, a bounded stretch of the lambda path, archives back to the host, all inside the timed region
  roofline     HBM roofline of the dominant kernel (saga_sparse_wave_kernel): algorithmic bytes per update
               (12*nnz_row + 8 + 4 + 8*K_y + 16*K = 1236 B) x updates per launch / kernel time
  cpu_baseline the reference's CPU path (oracle/_ref, else the restated oracle; g++ -O2, 1 thread: the reference is
               single-threaded) on a bounded sample of the same workload, timed on this box

One process per GPU. A single fit does not shard (the solver is a serial recurrence); with --gpus N every rank runs an
independent fit of the same shape (what cv folds / alpha grids are), no data-path collective: scaling = weak.
`--impl reference` times the CPU arm instead: oracle/_ref (the reference's own solver sources compiled against a stand-in
for Rcpp/Eigen, which are not in the image) when it was built, else the restated oracle.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # per-fit streams of batches (engine.cu); before CUDA starts

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NNZ_ROW = 100
OUT = sys.stdout
B_UPD = 12 * NNZ_ROW + 8 + 4 + 8 * 1 + 16 * 1          # 1236 B per sample-update


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", dest="n", type=int, default=1_000_000, help="n (rows of X); the metric is quoted at the default")
    ap.add_argument("--cols", dest="p", type=int, default=100_000, help="p (columns of X)")
    ap.add_argument("--lambda-ind", type=int, default=30)
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU time of the bounded oracle sample")
    ap.add_argument("--e2e-epochs", type=int, default=48,
                    help="epochs of the bounded sgdnet_fit_sparse call of the e2e leg (the whole 100-lambda path of this workload is 381)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


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


def control_for(lib, n_lambda=100, lambda_=None, maxit=1000):
    from sgdnet_b200 import api
    return api.build_control("binomial", 1, alpha=1.0, nlambda=n_lambda, lambda_min_ratio=1e-4, lambda_=lambda_,
                             maxit=maxit, standardize=False, intercept=True, thresh=1e-3, standardize_response=False,
                             debug=False)


def cpu_arm():
    """The CPU implementation that is timed: oracle/_ref (the reference's own src/sgdnet.cpp compiled against the
    Rcpp/Eigen stand-in, kind "reference") when it was built, else the restated oracle in its libm arithmetic (kind
    "port"). The two produce identical bits (tests/test_ref_cpu.py)."""
    from oracle_lib import load_oracle, load_reference_build
    ref = load_reference_build()
    if ref is not None:
        ref.lib.ref_set_force_debug(0)      # per-epoch loss passes are only needed for the tests' epoch counts
        return ref, "reference", ("reference src/sgdnet.cpp + saga-sparse.h compiled unmodified (g++ -O2, no FMA contraction) against "
                                  "the Rcpp/Eigen stand-in of oracle/refbuild (no R/Eigen in the image), 1 thread: the reference is single-threaded")
    oracle = load_oracle()
    oracle.lib.oracle_set_arith(0)      # std::exp/std::log, sequential sums: the most literal reading of the reference
    return oracle, "port", "CPU oracle restatement (libm arithmetic), g++ -O2, 1 thread: the reference is single-threaded"


def run_reference(args, rank, world):
    """CPU arm, single thread, `steps` epochs at the same lambda of the same workload."""
    if rank != 0:
        return
    from sgdnet_b200 import _abi
    oracle, kind, kind_note = cpu_arm()
    # bounded sample: the same generator at n rows capped so that warmup+steps epochs stay near --cpu-seconds
    est_rate = 0.4e6
    n = int(min(args.n, max(20_000, est_rate * args.cpu_seconds / max(1, args.steps + args.warmup))))
    x, y = make_workload(n, args.p)
    lam = path_lambda(oracle, x, y, args.lambda_ind)
    total = args.steps + args.warmup

    def run(epochs):
        ctl, keep = control_for(oracle, 1, [lam], maxit=epochs)
        ctl.tol = 0.0                                   # never converge early: exactly `epochs` epochs
        t0 = time.time()
        raw = oracle.fit(x, y.reshape(-1, 1), ctl, oracle.rng_from_seed(1))
        return raw, time.time() - t0

    # the timed steps are the LAST `steps` epochs of a (warmup + steps)-epoch fit: the same fit cut after `warmup`
    # epochs (same seed, same sequence) gives the time of the warm-up part, which is subtracted
    raw, wall = run(total)
    raw_w, _ = run(args.warmup) if args.warmup > 0 else (None, 0.0)
    solver_steps = raw.seconds_solver - (raw_w.seconds_solver if raw_w is not None else 0.0)
    ups = n * args.steps / solver_steps
    ms_step = solver_steps / args.steps * 1e3
    line = {
        "impl": "reference", "metric": "SAGA sample-updates/s", "value": ups, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, n_used=n),
        "cpu_baseline": {"value": ups, "unit": "updates/s", "cores": 1, "kind": kind,
                         "sample": f"last {args.steps} of {raw.npasses} epochs of n={n} rows (same generator, p={args.p}, {NNZ_ROW} nnz/row) at lambda[{args.lambda_ind}]; "
                                   + kind_note},
        "e2e": {"value": n * raw.npasses / wall, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


def path_lambda(lib, x, y, ind):
    """lambda[ind] of the automatic 100-lambda path of this workload (setup only: maxit=1, one lambda is enough to
    read lambda_max; the rest of the path is LogSpace)."""
    xc = x.tocsc()
    yb = (y - y.mean()) / y.std()
    lmax = y.std() * np.abs(xc.T @ yb).max() / x.shape[0]
    return float(np.exp(np.log(lmax) + ind * (np.log(lmax * 1e-4) - np.log(lmax)) / 99.0))


def config_dict(args, n_used=None):
    return {"workload": "BASELINE config 2: binomial lasso (alpha=1), sparse CSR 1M x 100k, 100 nnz/row, standardize=FALSE, "
                        "intercept=TRUE; step = one SAGA epoch at lambda[%d] of the 100-lambda path" % args.lambda_ind,
            "n": int(n_used if n_used is not None else args.n), "p": args.p, "nnz_row": NNZ_ROW, "family": "binomial",
            "penalty": "lasso", "fits_per_gpu": 1,
            "l2_policy": "inputs larger than L2: 1.2 GB CSR per pass vs 126 MB L2, random row order"}


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
    lib = sg.product()
    lib.check(lib.sym("set_device")(local), "set_device")

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
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        kernel_ms.append(ms.value)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    # device time of the timed region = sum of the CUDA-event brackets around the epoch kernels (library stream);
    # wall (host clock around barrier+sync) also contains index generation and upload for the next epoch
    dev_s = sum(kernel_ms) * 1e-3
    t = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall = float(t[0]), float(t[1])
    updates = n * args.steps * world
    value = updates / wall
    launches = 3 * args.steps   # per epoch: lag-scaling table (no-op for lasso), conflict codes, wavefront solver

    # ---------------- roofline of the dominant kernel (one launch = one epoch = n updates)
    peak, peak_src = peaks()
    per_launch_s = (sum(kernel_ms) / len(kernel_ms)) * 1e-3
    achieved = n * B_UPD / per_launch_s / 1e9
    traffic, traffic_src = None, None
    try:      # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on this workload
        with open(os.path.join(ROOT, "profiles", "r1_12_wave_fastpath_ncu_summary.json")) as fh:
            prof = json.load(fh)
        if n == prof["updates_per_launch"]:
            traffic, traffic_src = prof["dram_bytes_per_launch"], "profiles/r1_12_wave_fastpath_ncu_summary.json"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "saga_sparse_wave_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "bytes_per_update": B_UPD, "updates_per_launch": n, "launch_ms": per_launch_s * 1e3,
                "note": "serial recurrence: one CTA per fit, bounded by the intercept/gradient chain (about 500 cycles per update), not by HBM (DESIGN.md)"}
    lib.sym("session_destroy")(sess)

    # ---------------- e2e: sgdnet_fit_sparse with host buffers, bounded stretch of the path
    e2e = None
    if not args.no_e2e:
        lam = path_lambda(lib, x, y, args.lambda_ind)
        ne = max(2, args.e2e_epochs)
        ctl2, keep2 = control_for(lib, 1, [lam], maxit=ne)
        ctl2.tol = 0.0
        barrier()
        t0 = time.perf_counter()
        raw = lib.fit(m, ya, ctl2, lib.rng_from_seed(1 + rank))
        barrier()
        e2e_wall = time.perf_counter() - t0
        tt = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_wall = float(tt[0])
        h2d = m.i.nbytes + m.x.nbytes + 16 * n + ya.nbytes + 4 * n * raw.npasses
        d2h = raw.beta.nbytes + raw.a0.nbytes + raw.dev_ratio.nbytes
        e2e = {"value": n * raw.npasses * world / e2e_wall, "unit": "updates/s", "h2d_bytes_per_step": int(h2d / raw.npasses),
               "d2h_bytes_per_step": int(d2h / raw.npasses), "epochs": int(raw.npasses), "wall_s": e2e_wall,
               "setup_s": raw.seconds_setup, "solver_s": raw.seconds_solver,
               "note": "one sgdnet_fit_sparse call: host CSC in, CSR build + upload, %d epochs at one lambda, deviance, archives out" % raw.npasses}

    # ---------------- cpu baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        oracle, cpu_kind, cpu_note = cpu_arm()
        n_cpu = int(min(n, max(20_000, 0.4e6 * args.cpu_seconds / 3)))
        xs, ys = make_workload(n_cpu, args.p) if n_cpu != n else (x, y)
        lam = path_lambda(oracle, xs, ys, args.lambda_ind)
        def run_cpu(epochs):
            ctl3, keep3 = control_for(oracle, 1, [lam], maxit=epochs)
            ctl3.tol = 0.0
            return oracle.fit(xs, ys.reshape(-1, 1), ctl3, oracle.rng_from_seed(1))
        rawc, raww = run_cpu(3), run_cpu(1)             # epochs 2-3 are timed (epoch 1 runs on a cold cache)
        cpu_s = rawc.seconds_solver - raww.seconds_solver
        cpu = {"value": n_cpu * 2 / cpu_s, "unit": "updates/s", "cores": 1, "kind": cpu_kind,
               "sample": f"epochs 2-3 of a 3-epoch fit of n={n_cpu} rows (same generator, p={args.p}, {NNZ_ROW} nnz/row) at lambda[{args.lambda_ind}]; "
                         + cpu_note,
               "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {"metric": "SAGA sample-updates/s", "value": value, "unit": "updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args),
                "device_ms_per_step": dev_s / args.steps * 1e3, "value_device_only": updates / dev_s,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
