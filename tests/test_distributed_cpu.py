"""World-size-2 gloo test of the data-parallel protocol (no GPU): shard -> per-shard statistics -> ONE all-reduce of the
packed buffer -> finish on every rank.  The per-shard arithmetic is the CPU model of the device algorithm
(tests/algo_model.py); the sharding / reduction helpers are the product's (onoffgpf/parallel.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, N, M, D, ret):
    for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "zero-inflated-gp_b200"), ROOT):
        sys.path.insert(0, p)
    import algo_model as am
    from onoffgpf import parallel, synth
    from oracle import onoff_oracle as orc
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    X, Y = synth.make_data(N, D, 3)
    p = orc.as_params(**synth.make_params(M, M, D, 3))
    lo, hi = parallel.shard_bounds(N, rank, world)
    n_global = parallel.global_count(hi - lo)
    assert n_global == N
    num_data = 3 * N                                  # minibatch-style scale, identical on every rank
    scale = num_data / n_global
    like = {}

    def accumulate():
        st = am.accumulate(X[lo:hi], Y[lo:hi], p, scale)
        like.update(st)
        return am.pack_stats(st)

    def finish(buf):
        return am.finish(am.unpack_stats(buf, like), p, scale)

    e, dt, kl, g = parallel.evaluate_sharded(accumulate, finish)
    ret[rank] = (e, dt, kl, {k: torch.as_tensor(v).detach().reshape(-1).numpy() for k, v in g.items()})
    dist.destroy_process_group()


@pytest.mark.parametrize("N,M,D", [(301, 12, 2)])
def test_two_rank_gloo_matches_single_rank_and_oracle(N, M, D):
    import algo_model as am
    from onoffgpf import synth
    from oracle import onoff_oracle as orc
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, N, M, D, ret), nprocs=world, join=True)
    X, Y = synth.make_data(N, D, 3)
    p = orc.as_params(**synth.make_params(M, M, D, 3))
    e1, d1, k1, g1 = am.elbo_and_grads(X, Y, p, num_data=3 * N)
    eo, do, ko, go = orc.elbo_and_grads(X, Y, p, num_data=3 * N)
    for r in range(world):
        e, dt, kl, g = ret[r]
        assert e == pytest.approx(e1, rel=1e-13) and e == pytest.approx(eo, rel=1e-11)
        assert kl == pytest.approx(k1, rel=1e-14)            # KL enters once, not once per shard
        assert dt == pytest.approx(d1, rel=1e-13)
        for k in orc.PARAM_NAMES:
            ref = go[k].numpy().reshape(-1)
            assert np.linalg.norm(g[k] - ref) <= 1e-9 * np.linalg.norm(ref) + 1e-12, k
    # both ranks hold bitwise-identical results (same reduced buffer, same finish)
    assert ret[0][0] == ret[1][0]
    for k in orc.PARAM_NAMES:
        np.testing.assert_array_equal(ret[0][3][k], ret[1][3][k])


def _worker_kron(rank, world, port, N, Ms, Mt, ret):
    for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "zero-inflated-gp_b200"), ROOT):
        sys.path.insert(0, p)
    import algo_model_kron as amk
    from onoffgpf import parallel, synth
    from oracle import kron_oracle as kro
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    X, Y = synth.make_data(N, 3, 5)
    p = kro.as_params(**synth.make_kron_params(Ms, Mt, 2, 1, 5))
    lo, hi = parallel.shard_bounds(N, rank, world)
    scale = 2 * N / parallel.global_count(hi - lo)
    like = {}

    def accumulate():
        st, _ = amk.accumulate(X[lo:hi], Y[lo:hi], p, scale)
        like.update(st)
        return amk.pack_stats(st)

    def finish(buf):
        return amk.finish(amk.unpack_stats(buf, like), p, scale)

    e, dt, kl, g = parallel.evaluate_sharded(accumulate, finish)
    ret[rank] = (e, dt, kl, {k: torch.as_tensor(v).detach().reshape(-1).numpy() for k, v in g.items()})
    dist.destroy_process_group()


def test_two_rank_gloo_kronecker_variant():
    from onoffgpf import synth
    from oracle import kron_oracle as kro
    N, Ms, Mt = 203, 5, 7
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_kron, args=(world, port, N, Ms, Mt, ret), nprocs=world, join=True)
    X, Y = synth.make_data(N, 3, 5)
    p = kro.as_params(**synth.make_kron_params(Ms, Mt, 2, 1, 5))
    eo, do, ko, go = kro.elbo_and_grads(X, Y, p, num_data=2 * N)
    for r in range(world):
        e, dt, kl, g = ret[r]
        assert e == pytest.approx(eo, rel=1e-11) and kl == pytest.approx(ko, rel=1e-11) and dt == pytest.approx(do, rel=1e-11)
        for k in kro.PARAM_NAMES:
            ref = go[k].numpy().reshape(-1)
            assert np.linalg.norm(g[k] - ref) <= 1e-9 * np.linalg.norm(ref) + 1e-12, k
    assert ret[0][0] == ret[1][0]


def test_shard_bounds_cover_every_row_once():
    from onoffgpf.parallel import shard_bounds
    for N in (0, 1, 7, 1 << 24, (1 << 24) + 5):
        for W in (1, 2, 3, 4, 8):
            b = [shard_bounds(N, r, W) for r in range(W)]
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[i][1] == b[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)
