"""CPU, world_size 2, gloo: the multi-GPU host logic (frame ranges, max-over-ranks timing).  The data
path has no collective, so this is all there is to test without GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fs_uae_image_enhancer_project_b200.sharding import frame_range, max_over_ranks


def test_frame_ranges_partition_the_stream():
    for n in (0, 1, 7, 64, 4096, 4099):
        for world in (1, 2, 4, 8):
            rs = [frame_range(n, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    assert frame_range(4096, 8, 3) == (1536, 2048)          # BASELINE config 5: 512 frames per GPU
    with pytest.raises(ValueError):
        frame_range(10, 2, 2)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = frame_range(4096, world, rank)
    mine = torch.tensor([lo, hi], dtype=torch.int64)
    got = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(got, mine)
    covered = sorted((int(t[0]), int(t[1])) for t in got)
    slow = max_over_ranks(1.0 + rank)                     # rank 1 is the slow one
    dist.barrier()
    ret[rank] = (covered, slow)
    dist.destroy_process_group()


def test_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    for rank in (0, 1):
        covered, slow = ret[rank]
        assert covered == [(0, 2048), (2048, 4096)]
        assert slow == 2.0


def test_numa_binding_is_best_effort_without_a_gpu():
    """No GPU / no NVML / no sysfs topology: nothing is bound and nothing raises."""
    import os
    from fs_uae_image_enhancer_project_b200 import sharding
    before = os.sched_getaffinity(0)
    assert sharding.gpu_local_cpus(97) == []
    assert sharding.bind_to_gpu_numa_node(97) == []
    assert os.sched_getaffinity(0) == before
