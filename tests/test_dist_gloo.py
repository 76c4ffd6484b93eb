"""Multi-process host logic of the env-sharded render: world_size-2 `gloo` on CPU.
The render path itself has no collective; what is distributed is (a) which envs a rank renders and
(b) the optional gather of the observation tensors to rank 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sim_a_splat_b200 import dist as D
from sim_a_splat_b200 import scene as S


def test_shard_envs_partitions_contiguously():
    for E, G in ((256, 1), (256, 2), (2048, 8), (7, 2), (5, 8), (1, 4)):
        blocks = [D.shard_envs(E, G, r) for r in range(G)]
        flat = [e for b in blocks for e in b]
        assert flat == list(range(E))                      # disjoint, ordered, complete
        sizes = D.shard_sizes(E, G)
        assert sum(sizes) == E and max(sizes) - min(sizes) <= 1
        assert sizes == sorted(sizes, reverse=True)        # remainder goes to the first ranks


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, E: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    envs = D.shard_envs(E, world, rank)
    # each rank derives its own envs' link poses from rank-independent env seeds, so a 2-rank run renders
    # exactly the poses a 1-rank run would (what bench.py / VecSplatRenderer rely on)
    all_poses = S.random_link_poses(E, 3, seed=77)
    mine = all_poses[envs.start:envs.stop]
    # stand-in observation: deterministic function of (env id, pose) with the layout of camera_i = [E_local,3,H,W] u8
    obs = torch.from_numpy(np.stack([np.full((3, 4, 5), e % 251, dtype=np.uint8) for e in envs]) if len(envs) else np.zeros((0, 3, 4, 5), np.uint8))
    obs = obs + torch.from_numpy((np.abs(mine[:, 0, 9]) * 100).astype(np.uint8))[:, None, None, None]
    got = D.gather_observations(obs, dst=0)
    if rank == 0:
        assert got.shape == (E, 3, 4, 5) and got.dtype == torch.uint8
        want = np.stack([np.full((3, 4, 5), e % 251, dtype=np.uint8) for e in range(E)]) + (np.abs(all_poses[:, 0, 9]) * 100).astype(np.uint8)[:, None, None, None]
        np.testing.assert_array_equal(got.numpy(), want)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([1]))
    else:
        assert got is None
    # max-over-ranks timing reduction as bench.py does it
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == world
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("E", [8, 7])
def test_two_rank_gloo_shard_and_gather(tmp_path, E):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, E, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok.npy").exists()
