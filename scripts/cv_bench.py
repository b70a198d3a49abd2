"""BASELINE config 5 on one GPU: 10-fold cv_sgdnet x 5-alpha grid, binomial sparse (default n = 500k, p = 50k, 50 nnz/row),
every fit one CTA, all fits of a phase concurrent. Prints fits/s and aggregate sample-updates/s.
Usage: python scripts/cv_bench.py [n] [p] [nlambda] [maxit]"""
import json
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgdnet_b200 as sg
from sgdnet_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
nlambda = int(sys.argv[3]) if len(sys.argv) > 3 else 20
maxit = int(sys.argv[4]) if len(sys.argv) > 4 else 50
thresh = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-3
x, y = synth.binomial_sparse(n, p, 50, seed=1005)
perm = np.random.Generator(np.random.PCG64(1005)).permutation(n)
foldid = (perm % 10) + 1
lib = sg.product()
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
shard = None
if world > 1:      # one process per GPU (torchrun): fold fits dealt to the ranks, one all_gather of the score rows
    import torch
    import torch.distributed as dist
    from sgdnet_b200.shard import Shard
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib.check(lib.sym("set_device")(local), "set_device")
    shard = Shard.from_torch()
    dist.barrier()
t0 = time.perf_counter()
cv = sg.cv_sgdnet(x, y, family="binomial", alpha=[0.0, 0.25, 0.5, 0.75, 1.0], foldid=foldid, nlambda=nlambda,
                  standardize=False, maxit=maxit, thresh=thresh, seed=1000, backend=lib, shard=shard)
if world > 1:
    dist.barrier()
wall = time.perf_counter() - t0
mine = [f for f in cv.fold_fits if f is not None]
full = [f for f in cv.fits if f is not None and f.raw is not None]
upd_full = sum(int(f.npasses) * n for f in full)
upd_fold = sum(int(f.npasses) * f.nobs for f in mine)
solver_full = max([f.raw.seconds_solver for f in full] or [0.0])
solver_fold = max([f.raw.seconds_solver for f in mine] or [0.0])
print(json.dumps({"rank": rank, "full_fits": [(float(f.alpha), int(f.npasses), round(f.raw.seconds_solver, 2)) for f in full],
                  "fold_fits_npasses": [int(f.npasses) for f in mine],
                  "fold_fits_solver_s": [round(f.raw.seconds_solver, 2) for f in mine]}), file=sys.stderr)
if world > 1:
    dist.destroy_process_group()
if rank == 0:
  print(json.dumps({"n_gpus": world, "fold_fits_on_rank0": len(mine),"workload": f"cv_sgdnet 10 folds x 5 alphas, binomial sparse {n}x{p}, 50 nnz/row, nlambda={nlambda}, maxit={maxit}",
                  "fits": len(cv.fits) + len(cv.fold_fits), "wall_s": wall, "fits_per_s": (len(cv.fits) + len(cv.fold_fits)) / wall,
                  "thresh": thresh,
                  "updates_full_fits": upd_full, "updates_fold_fits": upd_fold,
                  "solver_s_full_phase": solver_full, "solver_s_fold_phase": solver_fold,
                  "agg_updates_per_s_fold_phase": upd_fold / solver_fold, "agg_updates_per_s_full_phase": upd_full / solver_full,
                  "alpha_min": cv.alpha_min, "lambda_min": cv.lambda_min}))
