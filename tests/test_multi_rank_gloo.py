"""N > 1 path of bench.py on CPU: two gloo ranks (world_size 2, 127.0.0.1).  The hot path shards by image with no
data-path collective ("replicas only"); the only collective is the MAX over ranks of the elapsed
time, from which rank 0 reports the whole-job throughput."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import bench

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = [10.0 + 5.0 * rank, 40.0 - 3.0 * rank]  # rank 0: (10, 40), rank 1: (15, 37)
    red = bench.max_over_ranks(local, dist, torch.device("cpu"))
    seeds = bench.image_seeds(rank, 3)
    gathered = [None] * world
    dist.all_gather_object(gathered, seeds)
    dist.barrier()
    if rank == 0:
        q.put((red, gathered, bench.aggregate_throughput(world, 50, red[0])))
    dist.destroy_process_group()


def test_two_rank_gloo_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    red, gathered, value = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert red == [15.0, 40.0]  # element-wise max over ranks
    assert set(gathered[0]).isdisjoint(gathered[1])  # replicas never share an image
    # whole-job images/s from the slowest rank's time: 2 ranks x 50 steps x IMGS_PER_GPU images per step
    assert abs(value - 2 * 50 * 2 / 0.015) < 1e-6


def test_single_rank_is_identity():
    sys.path.insert(0, ROOT)
    import bench

    assert bench.max_over_ranks([3.0, 4.0], None, torch.device("cpu")) == [3.0, 4.0]
    assert bench.aggregate_throughput(1, 10, 1000.0) == 10.0 * bench.IMGS_PER_GPU
    assert bench.aggregate_throughput(1, 10, 1000.0, images_per_step=1) == 10.0
