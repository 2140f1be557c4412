"""N>1 path on CPU: two gloo ranks run the multi-GPU plumbing of the tracking path (b200/multi.py) --
weight broadcast, sequence ownership, max-over-ranks timing, trajectory gather -- and each rank tracks its own
synthetic pair through the (host-emulated) device kernels; the results must equal the single-process ones."""
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
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _track_pair(seed):
    """E-matrix RANSAC + recoverPose on a seeded correspondence set through the C-ABI (hostsim build)."""
    import synthdata
    from b200 import tracking
    import importlib.util
    from b200 import native
    spec = importlib.util.spec_from_file_location("_hostsim_build", os.path.join(ROOT, "tests", "hostsim", "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    if os.path.join(ROOT, "tests", "hostsim") not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
    from runtime import HostsimRuntime
    rt = HostsimRuntime(native.Lib(m.build()))
    eng = tracking.Engine(376, 1241, rt)
    kp_ref, kp_cur, meta = synthdata.correspondences(seed, n=600)
    K = synthdata.kitti_intrinsics()
    rng = np.random.RandomState(seed)
    r = tracking.compute_pose_2d2d(eng, kp_ref, kp_cur, K, repeat=2, reproj_thre=0.2, rng=rng)
    return np.concatenate([r["R"].reshape(-1), r["t"].reshape(-1), [float(r["inliers"].sum())]])


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "df-vo_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import synthdata
    from b200 import multi
    # ---- weights: only rank 0 holds the real values before the broadcast
    flow_w = synthdata.liteflownet_weights()
    enc, dec = synthdata.monodepth2_weights(4869, 192, 640)
    want = multi.pack_weights([flow_w, enc, dec])[0].copy()
    if rank != 0:
        for d in (flow_w, enc, dec):
            for k in d:
                if hasattr(d[k], "shape"):
                    d[k] = np.zeros_like(d[k])
    flow_w, enc, dec = multi.broadcast_weights([flow_w, enc, dec], src=0)
    got = multi.pack_weights([flow_w, enc, dec])[0]
    assert np.array_equal(got, want), "rank %d: weights differ after broadcast" % rank
    # ---- ownership + per-rank tracking of the owned sequences
    mine = multi.rank_sequences(5, rank, world)
    poses = {s: _track_pair(100 + s) for s in mine[:1]}
    ms = multi.max_over_ranks(10.0 + rank)
    assert ms == 10.0 + world - 1
    allp = multi.gather_trajectories(poses, dst=0)
    # ---- pair-level sharding (batched many-frames mode): ragged contiguous blocks, one all_gather of per-pair statistics
    sh = multi.PairShard(5, rank, world)
    stats = np.stack([np.arange(sh.first, sh.first + sh.count, dtype=np.float64), 0.5 + np.arange(sh.first, sh.first + sh.count)], 1)
    allstats = multi.gather_pair_stats(stats, sh, world)
    assert allstats.shape == (5, 2) and np.array_equal(allstats[:, 0], np.arange(5.0)) and np.array_equal(allstats[:, 1], 0.5 + np.arange(5.0))
    if rank == 0:
        np.save(os.path.join(out_dir, "owned.npy"), np.array([multi.rank_sequences(5, r, world) for r in range(world)], dtype=object),
                allow_pickle=True)
        merged = {}
        for p in allp:
            merged.update(p)
        np.savez(os.path.join(out_dir, "poses.npz"), **{str(k): v for k, v in merged.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gloo(tmp_path, hostsim_lib):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    owned = np.load(tmp_path / "owned.npy", allow_pickle=True)
    flat = sorted(int(s) for row in owned for s in row)
    assert flat == list(range(5))                       # every sequence exactly one owner
    poses = np.load(tmp_path / "poses.npz")
    assert sorted(poses.files) == ["0", "1"]            # rank r tracked sequence r
    for k in poses.files:                               # same answer as a single process
        np.testing.assert_array_equal(poses[k], _track_pair(100 + int(k)))


def test_layout_roundtrip():
    from b200 import multi
    a = {"w": np.arange(6, dtype=np.float32).reshape(2, 3), "n": 3}
    b = {"b": np.ones(4, np.float32), "a": np.float32(2.0) * np.ones((1,), np.float32)}
    flat, table = multi.pack_weights([a, b])
    assert flat.size == 11 and [t[1] for t in table] == ["w", "a", "b"]
    a2, b2 = multi.unpack_weights(flat * 2, table, [dict(a), dict(b)])
    assert np.array_equal(a2["w"], a["w"] * 2) and a2["n"] == 3 and np.array_equal(b2["b"], 2 * b["b"])
    assert multi.rank_sequences(7, 1, 3) == [1, 4]
    # pair shards: contiguous, disjoint, covering, for every world size that does or does not divide the total
    for total, world in ((64, 1), (64, 2), (64, 8), (64, 3), (5, 8)):
        cover = []
        for r in range(world):
            sh = multi.PairShard(total, r, world)
            cover += list(range(sh.first, sh.first + sh.count))
            assert sh.owner_table()[r] == (sh.first, sh.count)
        assert cover == list(range(total))
    assert np.array_equal(multi.gather_pair_stats([[1.0, 2.0]], multi.PairShard(1, 0, 1), 1), [[1.0, 2.0]])
    assert multi.max_over_ranks(3.5) == 3.5
