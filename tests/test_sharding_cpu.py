"""World-size-2 gloo test of the multi-GPU host logic (sharding + the optional final gather).  The local
evaluator is stubbed with the CPU oracle here -- on GPUs it is HybridNLP.eval_batch; nothing else changes."""
import os
import socket

import numpy as np
import pytest

from quadruped_landing_b200.sharding import shard_bounds

torch = pytest.importorskip("torch")


def test_shard_bounds_tile_the_batch():
    for B in (0, 1, 7, 4096, 65537):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(B, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == B
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import quadruped_landing_b200 as ql
    from oracle.oracle import Oracle
    from quadruped_landing_b200.sharding import evaluate_sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = ql.build_problem(N=9, k_trans=4)
    o = Oracle(p)
    rng = np.random.default_rng(0)                       # same global batch on every rank
    Z = ql.initial_guess(p)[None, :] + 1e-2 * rng.standard_normal((13, p.n_nlp))

    def evaluate(Zl):
        r = o.eval_batch(Zl.numpy(), nthreads=1)
        return {k: torch.from_numpy(v) for k, v in r.items()}

    out, (lo, hi), f_all = evaluate_sharded(evaluate, torch.from_numpy(Z))
    ref = o.eval_batch(Z, nthreads=1)
    ok = (np.array_equal(f_all.numpy(), ref["f"]) and np.array_equal(out["jac"].numpy(), ref["jac"][lo:hi])
          and (lo, hi) == shard_bounds(13, world, rank))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
    assert res == [(0, True), (1, True)]
