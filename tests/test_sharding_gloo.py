"""Multi-process (world_size 2, gloo, CPU) test of the N > 1 path's host logic: block sharding of global replicate ids
and the final all-gather of the targets matrix.  The per-replicate work here is done by the oracle (this is a CPU
test); the GPU equivalent is test_run_graph_compaction_and_sharding_invariance."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from common import K, ROOT, load_oracle, make_case

N_TOTAL, WEEKS = 5, 12


def _targets_of(rid):
    case = make_case(300, 700 + rid, network_mode="surrogate")
    o = load_oracle(case, 1971, rid, t_cap=WEEKS + 3)
    for at in range(2, WEEKS + 2):
        o.step(at, K.M_ALL)
    return o.targets(WEEKS + 1, 6)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from xgc_msm_b200.shard import gather_targets, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    first, count = shard_range(N_TOTAL, rank, world)
    local = np.stack([_targets_of(first + k) for k in range(count)])
    full = gather_targets(local, N_TOTAL)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full)
    dist.destroy_process_group()


def test_shard_ranges_cover_every_replicate_once():
    from xgc_msm_b200.shard import shard_range
    for n, w in [(1000, 8), (5, 2), (7, 4), (3, 8)]:
        ids = [i for r in range(w) for i in range(shard_range(n, r, w)[0], sum(shard_range(n, r, w)))]
        assert ids == list(range(n))


def test_two_rank_gather_equals_single_process(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ref = np.stack([_targets_of(r) for r in range(N_TOTAL)])
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npy")
        assert got.shape == ref.shape and np.array_equal(got.view(np.uint64), ref.view(np.uint64))
