"""N > 1 host path on CPU: two gloo ranks shard the walkers, each runs the (emulated) kernels on its shard,
log-likelihoods are all-gathered and must equal the single-process evaluation."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_indices_round_robin():
    import py21cmmc_b200  # noqa: F401
    from py21cmmc_b200.walkers import shard_indices

    assert shard_indices(7, 0, 2) == [0, 2, 4, 6]
    assert shard_indices(7, 1, 2) == [1, 3, 5]
    assert sorted(sum((shard_indices(64, r, 8) for r in range(8)), [])) == list(range(64))


def _worker(rank, world, port, nwalkers, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from tests.emu.build_emu import build
        import py21cmmc_b200  # noqa: F401
        from py21cmmc_b200 import _capi, chain as chain_mod, core as core_mod, likelihood as lk_mod
        from py21cmmc_b200.walkers import BatchedWalkerPool
        from tests.helpers import synth_box

        _capi._testing_set_library(build())
        shape, box_len = (12, 12, 36), 18.0
        up = core_mod.UserParams(BOX_LEN=box_len, HII_DIM=shape[0])
        # the "simulator": parameter p[0] selects the synthetic lightcone
        c = core_mod.CoreLightConeModule(lambda params: synth_box(shape, int(round(params.seed)) if len(params) else 0), up)
        lk = lk_mod.Likelihood1DPowerLightcone(nchunks=3, n_psbins=6, simulate=True, min_k=0.3, max_k=4.0)
        params = chain_mod.Params(("seed", [0.0, 0.0, 100.0, 1.0]))
        ch = chain_mod.build_computation_chain(c, lk, params=params)
        plist = [[float(s)] for s in range(1, nwalkers + 1)]
        pool = BatchedWalkerPool(ch, rank, world)
        res = pool.map(None, plist)
        q.put((rank, [r[0] for r in res], [r[1] is not None for r in res]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_process():
    import torch.multiprocessing as mp

    from tests.emu.build_emu import build

    build()
    nwalkers, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, nwalkers, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, lnl, has_blob = q.get(timeout=120)
        got[rank] = (lnl, has_blob)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank holds every walker's lnL, identical bits on both ranks
    assert got[0][0] == got[1][0]
    assert got[0][1] == [True, False, True, False, True] and got[1][1] == [False, True, False, True, False]
    # single-process reference through the same classes
    q1 = ctx.Queue()
    p = ctx.Process(target=_worker, args=(0, 1, port + 1, nwalkers, q1))
    p.start()
    _, lnl1, _ = q1.get(timeout=120)
    p.join(timeout=60)
    assert np.array_equal(np.array(lnl1), np.array(got[0][0]))
    assert all(np.isfinite(lnl1)) and len(set(lnl1)) == nwalkers


def _nccl_worker(rank, world, port, nwalkers, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        import py21cmmc_b200  # noqa: F401
        from py21cmmc_b200 import chain as chain_mod, core as core_mod, likelihood as lk_mod
        from py21cmmc_b200.walkers import BatchedWalkerPool
        from tests.helpers import synth_box

        shape, box_len = (128, 128, 512), 192.0   # two 256-slice chunks: the fused kernels
        up = core_mod.UserParams(BOX_LEN=box_len, HII_DIM=shape[0])
        c = core_mod.CoreLightConeModule(lambda params: synth_box(shape, int(round(params.seed)) if len(params) else 0), up)
        lk = lk_mod.Likelihood1DPowerLightcone(nchunks=2, n_psbins=12, simulate=True, min_k=0.1, max_k=2.0, device=rank,
                                               edge_rule="legacy_1.001")
        ch = chain_mod.build_computation_chain(c, lk, params=chain_mod.Params(("seed", [0.0, 0.0, 100.0, 1.0])))
        plist = [[float(s)] for s in range(1, nwalkers + 1)]
        pool = BatchedWalkerPool(ch, rank, world, device=torch.device("cuda", rank))
        res = pool.map(None, plist)
        half = pool.map(None, plist[: nwalkers // 2])     # a half-ensemble batch (ensemble.py:166-170)
        q.put((rank, [r[0] for r in res], [r[1] is not None for r in res], [r[0] for r in half]))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_nccl_sharding_matches_single_gpu():
    """BatchedWalkerPool over NCCL on two B200s: every rank ends with every walker's lnL, bit-identical to one GPU
    evaluating all walkers (VERDICT r1 missing #7).  Needs two GPUs: `gpurun --gpus 2`."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    nwalkers, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, nwalkers, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, lnl, has_blob, half = q.get(timeout=300)
        got[rank] = (lnl, has_blob, half)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got[0][0] == got[1][0] and got[0][2] == got[1][2] == got[0][0][: nwalkers // 2]
    assert got[0][1] == [True, False, True, False, True] and got[1][1] == [False, True, False, True, False]
    q1 = ctx.Queue()
    p = ctx.Process(target=_nccl_worker, args=(0, 1, port + 1, nwalkers, q1))
    p.start()
    _, lnl1, _, _ = q1.get(timeout=300)
    p.join(timeout=120)
    assert lnl1 == got[0][0] and all(np.isfinite(lnl1)) and len(set(lnl1)) == nwalkers
