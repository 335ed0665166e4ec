"""world_size-2 `gloo` test of the N>1 host logic (no GPU needed): contiguous
sharding, per-rank generation of its own sub-range of the global frame stream,
and the one collective on the path — the int64 statistics all-reduce — must
reproduce the single-process result exactly.  The plans themselves come from
the CPU oracle here (tests may use it); on the GPU box the same flow runs with
pp_plan_batch + pp_stats_batch + NCCL (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def numpy_stats(abi, plans):
    """Mirror of stats_kernel (csrc/pp_plan.cu) for host-side plans."""
    st = np.zeros(abi.STATS_LEN, dtype=np.int64)
    st[abi.STAT_FRAMES] = plans.n
    st[abi.STAT_POINTS] = plans.n_points.sum()
    st[abi.STAT_TARGET_LANE0:abi.STAT_TARGET_LANE0 + 3] = np.bincount(plans.target_lane, minlength=3)
    st[abi.STAT_EGO_LANE0:abi.STAT_EGO_LANE0 + 3] = np.bincount(plans.ego_lane, minlength=3)
    st[abi.STAT_LANE_CHANGES] = (plans.target_lane != plans.ego_lane).sum()
    for b in range(abi.NUM_FLAGS):
        st[abi.STAT_FLAG0 + b] = ((plans.flags >> b) & 1).sum()
    x, y = plans.next_x, plans.next_y
    ok = np.isfinite(x) & np.isfinite(y) & (np.abs(x) < 1e12) & (np.abs(y) < 1e12)
    ok &= np.arange(x.shape[1])[None, :] < plans.n_points[:, None]
    st[abi.STAT_XSUM] = (np.trunc(np.where(ok, x, 0) * 256.0).astype(np.int64).sum()
                         + np.trunc(np.where(ok, y, 0) * 256.0).astype(np.int64).sum())
    return st


def numpy_fstats(abi, plans):
    """Mirror of pp_fstats_batch (definition in include/pp.h) for host-side plans."""
    x, y, npts = plans.next_x, plans.next_y, plans.n_points
    out = np.array([np.inf] * abi.FSTAT_NMIN + [-np.inf] * (abi.FSTATS_LEN - abi.FSTAT_NMIN))

    def upd(i, vals):
        vals = vals[np.isfinite(vals)]
        if len(vals):
            out[i] = min(out[i], vals.min()) if i < abi.FSTAT_NMIN else max(out[i], vals.max())
    upd(0, plans.ego_speed), upd(3, plans.ego_speed)
    upd(1, plans.target_speed), upd(4, plans.target_speed)
    with np.errstate(invalid="ignore", over="ignore"):
        vx, vy = (x[:, 1:] - x[:, :-1]) * 50, (y[:, 1:] - y[:, :-1]) * 50
        sp = np.sqrt(vx * vx + vy * vy)
        k = np.arange(sp.shape[1])[None, :]
        live = sp[k < (npts[:, None] - 1)]
        upd(2, live), upd(5, live)
        ax, ay = (vx[:, 1:] - vx[:, :-1]) * 50, (vy[:, 1:] - vy[:, :-1]) * 50
        acc = np.sqrt(ax * ax + ay * ay)
        k = np.arange(acc.shape[1])[None, :]
        upd(6, acc[k < (npts[:, None] - 2)])
    return out


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    import checkers
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pp = checkers.load_pkg()
    from carnd_path_planning_project_b200 import parallel
    lo, hi = parallel.shard_range(n_total, rank, world)
    m = pp.Map()
    frames = pp.synth_frames(m, hi - lo, 12, seed=99, first_frame=lo, rare_permille=100)
    plans = checkers.Checker("oracle").plan(frames, cars=False)
    st = torch.from_numpy(numpy_stats(checkers.abi, plans))
    parallel.allreduce_stats(st)
    fs = torch.from_numpy(numpy_fstats(checkers.abi, plans))
    parallel.allreduce_fstats(fs, checkers.abi.FSTAT_NMIN)  # what pp_stats_reduce does on GPUs
    q.put((rank, lo, hi, st.numpy().copy(), plans.target_lane.copy(), fs.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_reduce_to_the_single_rank_result(pp, pmap, oracle, abi):
    import torch.multiprocessing as mp
    n_total, world = 3001, 2  # odd on purpose: uneven shards
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole_frames = pp.synth_frames(pmap, n_total, 12, seed=99, rare_permille=100)
    whole = oracle.plan(whole_frames, cars=False)
    want = numpy_stats(abi, whole)
    assert results[0][1:3] == (0, 1500) and results[1][1:3] == (1500, 3001)
    want_f = numpy_fstats(abi, whole)
    for r in results:
        assert np.array_equal(r[3], want)  # every rank holds the global sums
        assert np.array_equal(r[5], want_f)  # ... and the global minima / maxima, bit for bit
    assert np.isfinite(want_f).all() and want_f[abi.FSTAT_NMIN + 3] > 0
    assert np.array_equal(np.concatenate([r[4] for r in results]), whole.target_lane)
    assert want[abi.STAT_FRAMES] == n_total and want[abi.STAT_POINTS] > 0


def test_shard_range_partitions_exactly(pp):
    from carnd_path_planning_project_b200 import parallel
    for n in (0, 1, 7, 1 << 20, 64_000_001):
        for g in (1, 2, 4, 8):
            edges = [parallel.shard_range(n, r, g) for r in range(g)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(g - 1))
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)


def test_merge_fstats_rule(pp, abi):
    from carnd_path_planning_project_b200 import parallel
    a = np.array([1.0, 2.0, 3.0, 10.0, 20.0, 30.0, 40.0])
    b = np.array([0.5, 9.0, 3.0, 11.0, 5.0, 30.0, 41.0])
    got = parallel.merge_fstats(a, b, abi.FSTAT_NMIN)
    assert list(got) == [0.5, 2.0, 3.0, 11.0, 20.0, 30.0, 41.0]
    ident = np.array([np.inf] * abi.FSTAT_NMIN + [-np.inf] * (abi.FSTATS_LEN - abi.FSTAT_NMIN))
    assert np.array_equal(parallel.merge_fstats(a, ident, abi.FSTAT_NMIN), a)
