"""CPU, world_size 2 over gloo: the host logic of the data-parallel path.

Each rank takes its contiguous row shard (hetmogp_b200.shard_rows), computes the per-shard sufficient statistics --
here with the CPU oracle standing in for the CUDA engine, which needs a GPU -- packs them into ONE flat fp64 buffer
and sum-all-reduces it (the single collective of a step); the reduced statistics and the ELBO assembled from them
must equal the unsharded evaluation.  The same pack -> all_reduce -> finish sequence runs on NCCL in
Engine.evaluate(group=...) (GPU: tests/test_gpu_parity.py::test_shard_sum_equals_whole_large_n, bench.py --gpus N).
"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(o):
    return np.concatenate([o["VE_sum"].ravel(), o["sdv"].ravel(), o["sma"].ravel(), o["svc"].ravel(),
                           o["dVE_dmu"].ravel(), o["dVE_dS"].ravel()])


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from hetmogp_b200 import shard_rows, synth
    from oracle import diag_oracle
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    prob = synth.make_problem([("Gaussian", 0.5), ("Bernoulli",), ("Categorical", 3)], [301, 57, 1], 12, 2, seed=4)
    N = [x.shape[0] for x in prob["X"]]
    begin, count = shard_rows(N, rank, world)
    sl = [slice(b, b + c) for b, c in zip(begin, count)]
    part = diag_oracle.elbo_and_grads(prob, row_slices=sl, want_hyper=False)
    buf = torch.from_numpy(_pack(part))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)          # the ONE collective of a step
    whole = diag_oracle.elbo_and_grads(prob, want_hyper=False)
    ref = _pack(whole)
    err = float(np.max(np.abs(buf.numpy() - ref)) / np.max(np.abs(ref)))
    T = len(N)
    elbo = float(buf.numpy()[:T].sum() - whole["KL"])     # KL is replicated, not sharded
    q.put((rank, err, elbo, float(whole["log_marginal"][0, 0]), count))
    dist.destroy_process_group()


def test_two_rank_allreduce_identity():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    counts = [r[4] for r in sorted(res)]
    assert [a + b for a, b in zip(*counts)] == [301, 57, 1]          # shards partition every task, incl. the 1-row task
    for rank, err, elbo, ref, _ in res:
        assert err < 1e-12, (rank, err)
        assert abs(elbo - ref) < 1e-10 * abs(ref)
