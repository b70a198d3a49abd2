"""Multi-rank path of cv_sgdnet on the CPU: world size 2 over gloo, CPU oracle as the backend (the product needs a
GPU). Checks the fit-to-rank assignment and that the sharded run equals the single-process run after ONE all_gather
of the per-fit score rows."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from sgdnet_b200.shard import Shard, assign_longest_first


def test_longest_first_assignment_is_a_balanced_partition():
    costs = [500, 500, 500, 50, 50, 50, 50, 50, 50, 40, 30, 20]
    for world in (1, 2, 3, 4, 8):
        parts = assign_longest_first(costs, world)
        assert sorted(k for part in parts for k in part) == list(range(len(costs)))
        loads = [sum(costs[k] for k in part) for part in parts]
        assert max(loads) - min(loads) <= max(costs)
    assert assign_longest_first(costs, 2) == assign_longest_first(list(costs), 2)      # deterministic


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker_batched(rank, world, port, out_dir):
    """the batched path (one sgdnet_fit_batch_* call per rank): full-data fits are sharded too, a rank asks for the
    lambda path only (`path_only`) of the alphas whose full fit another rank owns, and the selected alpha's fit is
    broadcast from its owner."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sgdnet_b200 as sg
    import synth
    from oracle_lib import load_oracle
    oracle = load_oracle()
    x, y = synth.binomial_sparse(600, 80, 8, seed=77)
    foldid = (np.random.Generator(np.random.PCG64(3)).permutation(600) % 4) + 1
    sh = Shard.from_torch()
    cv = sg.cv_sgdnet(x, y, family="binomial", alpha=[0.0, 0.5, 1.0], foldid=foldid, nlambda=6, standardize=False,
                      maxit=60, seed=500, backend=oracle, shard=sh)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), summary=cv.cv_summary, raw=np.stack(cv.cv_raw),
             mine=np.array([k for k, f in enumerate(cv.fold_fits) if f is not None]),
             full=np.array([k for k, f in enumerate(cv.fits) if f is not None]),
             best=np.array([cv.alpha_min, cv.lambda_min, cv.lambda_1se]), best_beta=cv.fit.beta, best_a0=cv.fit.a0,
             lambdas=np.stack(cv.lambda_))
    dist.barrier()
    dist.destroy_process_group()


def test_cv_batched_and_sharded_with_full_fits_dealt_to_the_ranks(tmp_path, oracle):
    import torch.multiprocessing as mp
    import sgdnet_b200 as sg
    import synth
    world, port = 2, _free_port()
    mp.start_processes(_worker_batched, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(world))
    assert sorted(list(r0["mine"]) + list(r1["mine"])) == list(range(12))
    # every full-data fit has one owner; the selected alpha's fit is on both ranks after the broadcast, nothing else is
    both = set(r0["full"]) & set(r1["full"])
    assert set(r0["full"]) | set(r1["full"]) == {0, 1, 2} and len(both) == 1
    assert len(r0["full"]) < 3 or len(r1["full"]) < 3
    for f in ("raw", "summary", "best", "best_beta", "best_a0", "lambdas"):
        np.testing.assert_array_equal(r0[f], r1[f], err_msg=f)
    x, y = synth.binomial_sparse(600, 80, 8, seed=77)
    foldid = (np.random.Generator(np.random.PCG64(3)).permutation(600) % 4) + 1
    ref = sg.cv_sgdnet(x, y, family="binomial", alpha=[0.0, 0.5, 1.0], foldid=foldid, nlambda=6, standardize=False,
                       maxit=60, seed=500, backend=oracle, batched=False)
    np.testing.assert_array_equal(np.stack(ref.cv_raw), r0["raw"])
    np.testing.assert_array_equal(ref.cv_summary, r0["summary"])
    np.testing.assert_array_equal(ref.fit.beta, r0["best_beta"])
    np.testing.assert_array_equal(ref.fit.a0, r0["best_a0"])
    np.testing.assert_array_equal(np.stack(ref.lambda_), r0["lambdas"])


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sgdnet_b200 as sg
    import synth
    from oracle_lib import load_oracle
    oracle = load_oracle()
    x, y = synth.binomial_sparse(600, 80, 8, seed=77)
    foldid = (np.random.Generator(np.random.PCG64(3)).permutation(600) % 4) + 1
    sh = Shard.from_torch()
    assert (sh.rank, sh.world) == (rank, world)
    cv = sg.cv_sgdnet(x, y, family="binomial", alpha=[0.0, 0.5, 1.0], foldid=foldid, nlambda=6, standardize=False,
                      maxit=60, seed=500, backend=oracle, batched=False, shard=sh)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), summary=cv.cv_summary, raw=np.stack(cv.cv_raw),
             mine=np.array([k for k, f in enumerate(cv.fold_fits) if f is not None]),
             best=np.array([cv.alpha_min, cv.lambda_min, cv.lambda_1se]))
    dist.barrier()
    dist.destroy_process_group()


def test_cv_sharded_over_two_gloo_ranks_matches_single_process(tmp_path, oracle):
    import torch.multiprocessing as mp
    import sgdnet_b200 as sg
    import synth
    world, port = 2, _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(world))
    # the two ranks hold disjoint halves of the 12 fold fits and agree on everything gathered
    assert sorted(list(r0["mine"]) + list(r1["mine"])) == list(range(12))
    assert len(r0["mine"]) == len(r1["mine"]) == 6
    np.testing.assert_array_equal(r0["raw"], r1["raw"])
    np.testing.assert_array_equal(r0["summary"], r1["summary"])
    # and equal the unsharded run bit for bit
    x, y = synth.binomial_sparse(600, 80, 8, seed=77)
    foldid = (np.random.Generator(np.random.PCG64(3)).permutation(600) % 4) + 1
    ref = sg.cv_sgdnet(x, y, family="binomial", alpha=[0.0, 0.5, 1.0], foldid=foldid, nlambda=6, standardize=False,
                       maxit=60, seed=500, backend=oracle, batched=False)
    np.testing.assert_array_equal(np.stack(ref.cv_raw), r0["raw"])
    np.testing.assert_array_equal(ref.cv_summary, r0["summary"])
    np.testing.assert_array_equal([ref.alpha_min, ref.lambda_min, ref.lambda_1se], r0["best"])
    assert not np.isnan(r0["raw"]).any()
