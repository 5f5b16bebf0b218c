"""Worker launched by tests/test_multiproc.py under torch.distributed.run (one process per rank).

  mode gloo : CPU.  Host-side logic of the N>1 path: chain sharding, the sum-allreduce used by the
              library's callback, and that the two-phase pooled merge over ranks equals the global merge
              (per-chain accumulators come from the oracle -- test infrastructure).
  mode nccl : one GPU per rank.  Sharded run (no collective) equals the single-handle run bit for bit;
              pooled adaptation and R-hat/ESS over NCCL agree with a single handle that owns every chain.
Prints "WORKER_OK <rank>" on success.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import mcmcf90_b200 as mb
from mcmcf90_b200 import parallel
from tests import cases

NTOT = 50  # not divisible by 4: ragged shards


def two_phase_merge(w, mean, cmat):
    """Host mirror of pool.cuh's two phases with an allreduce between them."""
    d = mean.shape[1]
    use = w > 0
    b1 = np.zeros(1 + d)
    b1[0] = w[use].sum()
    b1[1:] = (w[use, None] * mean[use]).sum(0)
    parallel.allreduce_host(b1)
    mu = b1[1:] / b1[0]
    dm = mean[use] - mu
    s2 = ((w[use] - 1.0)[:, None, None] * cmat[use] + w[use, None, None] * dm[:, :, None] * dm[:, None, :]).sum(0)
    parallel.allreduce_host(s2)
    return b1[0], mu, s2 / (b1[0] - 1.0)


def run_gloo():
    from oracle import oracle as O
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert parallel.env_rank()[:2] == (rank, world)
    # shards tile [0, NTOT) without gaps or overlap
    n, off = parallel.shard(NTOT, rank, world)
    spans = [None] * world
    dist.all_gather_object(spans, (n, off))
    assert spans[0][1] == 0 and sum(s[0] for s in spans) == NTOT
    for a, b in zip(spans[:-1], spans[1:]):
        assert a[1] + a[0] == b[1]
    assert max(s[0] for s in spans) - min(s[0] for s in spans) <= 1
    # allreduce on host buffers
    v = np.arange(5.0) + rank
    parallel.allreduce_host(v)
    assert np.array_equal(v, world * np.arange(5.0) + sum(range(world)))
    # pooled merge over ranks == global merge; accumulators of this rank's chains from the oracle
    blob = O.blob_expreg(cases.DATA_X, cases.DATA_Y)
    nml = dict(cases.NML_DRAM, nsimu=301)
    par0 = cases.PAR0 * (1 + 0.01 * np.random.default_rng(1).normal(size=(NTOT, 2)))
    out = O.run_batch(O.make_cfg(**nml), O.MODEL_EXPREG, blob, par0[off:off + n], cases.CMAT0, cases.SIGMA2, cases.NOBS,
                      seed=9, chain0=off, nthreads=2)
    w = np.full(n, 301.0)
    cm = out["cmat"]
    cm = np.triu(cm.transpose(0, 2, 1)) + np.transpose(np.triu(cm.transpose(0, 2, 1), 1), (0, 2, 1))
    W, mu, cov = two_phase_merge(w, out["mean"], cm)
    allw, allm, allc = [None] * world, [None] * world, [None] * world
    dist.all_gather_object(allw, w); dist.all_gather_object(allm, out["mean"]); dist.all_gather_object(allc, cm)
    Wg, mug, covg = O.pooled_moments(np.concatenate(allw), np.concatenate(allm), np.concatenate(allc))
    assert W == Wg
    np.testing.assert_allclose(mu, mug, rtol=1e-13)
    np.testing.assert_allclose(cov, covg, rtol=1e-10)
    # the global chain id keys the stream: this rank's chains equal the same chains of a one-rank run
    if rank == 0:
        full = O.run_batch(O.make_cfg(**nml), O.MODEL_EXPREG, blob, par0, cases.CMAT0, cases.SIGMA2, cases.NOBS, seed=9,
                           chain0=0, nthreads=2)
        assert np.array_equal(full["par"][off:off + n], out["par"])
    dist.barrier()
    dist.destroy_process_group()


def make_sampler(n, off, par0, dev, **kw):
    blob = mb.models.blob_expreg(cases.DATA_X, cases.DATA_Y)
    nml = dict(cases.NML_DRAM, nsimu=401, adaptint=50)
    s = mb.Sampler(mb.default_config(nchains=n, chain_offset=off, seed=11, device=dev, **nml, **kw))
    s.set_data(blob)
    s.set_initial(par0[off:off + n], cases.CMAT0, cases.SIGMA2, cases.NOBS)
    return s


def run_nccl():
    rank, world, local = parallel.env_rank()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ntot = 4096 + 10
    n, off = parallel.shard(ntot, rank, world)
    par0 = cases.PAR0 * (1 + 0.01 * np.random.default_rng(1).normal(size=(ntot, 2)))

    # (1) no collective: shard == the same chains of a single handle, bit for bit
    s = make_sampler(n, off, par0, local)
    s.run(400)
    mine = (s.fetch("par"), s.fetch("cmat"), s.fetch("counters"))
    s.close()
    if rank == 0:
        s1 = make_sampler(ntot, 0, par0, local)
        s1.run(400)
        ref = (s1.fetch("par"), s1.fetch("cmat"), s1.fetch("counters"))
        s1.close()
        for a, b in zip(mine, ref):
            assert np.array_equal(a, b[off:off + n])

    # (2) pooled adaptation + diagnostics over NCCL == one handle owning every chain (rounding level)
    kw = dict(pool_adapt=1, diag_stride=10, diag_lags=4)
    s = make_sampler(n, off, par0, local, **kw)
    assert parallel.attach(s) == (world > 1)
    s.run(400)
    W, mu, cov = s.pool_fetch()
    dg = s.diagnostics()
    cnt = s.counters()
    s.close()
    assert W == ntot * 401.0 and dg["nchains"] == ntot and dg["nsnap"] == 40
    res = [None] * world
    dist.all_gather_object(res, (W, mu, cov, dg["rhat"], dg["ess"], cnt["stayed"]))
    for r in res[1:]:  # every rank holds the same pooled statistics
        assert r[0] == res[0][0] and np.array_equal(r[1], res[0][1]) and np.array_equal(r[2], res[0][2])
        assert np.array_equal(r[3], res[0][3]) and np.array_equal(r[4], res[0][4])
    if rank == 0:
        s1 = make_sampler(ntot, 0, par0, local, **kw)
        s1.run(400)
        W1, mu1, cov1 = s1.pool_fetch()
        dg1 = s1.diagnostics()
        st1 = s1.counters()["stayed"]
        s1.close()
        assert W1 == W
        np.testing.assert_allclose(mu, mu1, rtol=1e-12)
        np.testing.assert_allclose(cov, cov1, rtol=1e-9)
        np.testing.assert_allclose(dg["rhat"], dg1["rhat"], rtol=1e-6)
        np.testing.assert_allclose(dg["ess"], dg1["ess"], rtol=1e-3)
        stayed = np.concatenate([r[5] for r in res])
        assert (stayed != st1).mean() < 0.02  # rounding-level differences in the pooled factor flip few chains
        assert (dg["rhat"] < 1.05).all()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    {"gloo": run_gloo, "nccl": run_nccl}[sys.argv[1]]()
    sys.stdout.write("WORKER_OK %s\n" % os.environ.get("RANK", "0"))
    sys.stdout.flush()
