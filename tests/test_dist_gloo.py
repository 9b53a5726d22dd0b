"""world_size-2 gloo test of the data-parallel host logic (no GPU): the sum all-reduce of
per-shard mean gradients scaled by 1/world equals the single-process gradient of the global
batch, parameters broadcast from rank 0, and epoch scalars agree on every rank."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dca_oracle as O
from tests.util import synth_counts


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dca_b200 import dist as D
    B, G = 32, 24
    Y = synth_counts(B, G, 5); X, sf = O.normalize_inputs(Y)
    p0 = O.init_params(G, G, (8, 4, 8), "zinb-conddisp", False, seed=3 + rank, dtype=np.float64)  # differ per rank
    names = sorted(p0)
    flat = torch.from_numpy(np.concatenate([p0[k].reshape(-1) for k in names]))
    D.broadcast_(flat, src=0)                                         # rank 0 wins
    off = 0
    for k in names:
        n = p0[k].size; p0[k] = flat[off:off + n].numpy().reshape(p0[k].shape).copy(); off += n
    net = O.OracleNet(G, G, (8, 4, 8), "zinb-conddisp", False, dtype=np.float64, params=p0)
    lo, hi = D.shard_bounds(B, rank, world, equal=True)
    loss, g = net.loss_and_grads(X[lo:hi].astype(np.float64), Y[lo:hi].astype(np.float64), sf[lo:hi].astype(np.float64))
    gflat = torch.from_numpy(np.concatenate([g[k].reshape(-1) for k in sorted(g)] + [np.array([loss])]))
    D.all_reduce_sum_(gflat); gflat *= 1.0 / world
    acc = D.all_reduce_sum_host(np.array([loss * (hi - lo), hi - lo], np.float64), "cpu")
    if rank == 0:
        q.put((gflat.numpy(), acc, {k: v for k, v in p0.items()}))
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_global_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    gflat, acc, p0 = q.get(timeout=120)
    for p in procs: p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    B, G = 32, 24
    Y = synth_counts(B, G, 5); X, sf = O.normalize_inputs(Y)
    net = O.OracleNet(G, G, (8, 4, 8), "zinb-conddisp", False, dtype=np.float64, params=p0)
    loss, g = net.loss_and_grads(X.astype(np.float64), Y.astype(np.float64), sf.astype(np.float64))
    ref = np.concatenate([g[k].reshape(-1) for k in sorted(g)] + [np.array([loss])])
    np.testing.assert_allclose(gflat, ref, rtol=1e-9, atol=1e-12)
    assert acc[1] == B and abs(acc[0] / acc[1] - loss) < 1e-9
