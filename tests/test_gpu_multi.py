"""GPU, world size 2 over NCCL (skipped on a single-GPU box; run with `gpurun --gpus 2`):
strip sharding of one frame. Rank r renders its rows on GPU r; the frame is assembled on rank 0
(a) by NCCL send/recv of finished strips and (b) with the tile stores going straight into rank 0's
framebuffer through peer memory. Both must be byte-identical to the single-GPU frame."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import ctypes as C
    import torch
    import torch.distributed as dist
    import minirender_b200 as m
    from minirender_b200 import cabi, scenes, sharding
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    lib = cabi.load()
    be = m.Backend()
    setup = scenes.bench_scene(be, width=1280, height=720, objects=6, m=60, n=60, usetex=True)
    r = setup.apply(m.Renderer(be))
    r.set_device(rank)
    ctx = r.context_ptr()
    h, w = setup.height, setup.width
    rb, re = sharding.strip_rows(h, rank, world)
    results = {}
    # (a) render own strip, gather with NCCL send/recv into rank 0's framebuffer
    r.clear()
    r.set_row_range(rb, re)
    r.render()
    r.synchronize()
    img, dep = sharding.device_tensors(lib, ctx, h, w, torch.device("cuda", rank))
    sharding.gather_strips(img, dep, h, rank, world, dist, dst=0)
    torch.cuda.synchronize()
    if rank == 0:
        results["gather_image"], results["gather_depth"] = img.cpu().numpy(), dep.cpu().numpy()
    dist.barrier()
    # (b) peer target: stores of every rank land in rank 0's buffers over NVLink, no gather step
    r.set_row_range(0, 0)
    r.set_background((0.3, 0.3, 0.3))
    r.clear()
    r.set_background(setup.background)
    r.synchronize()
    dist.barrier()
    close = sharding.open_peer_target(lib, ctx, rank, world, dist, dst=0)
    r.set_row_range(rb, re)
    r.render()
    r.synchronize()
    dist.barrier()
    close()
    if rank == 0:
        results["peer_image"], results["peer_depth"] = r.get_image().copy(), r.get_depth().copy()
        # single-GPU reference frame
        r.set_row_range(0, 0)
        r.render()
        results["full_image"], results["full_depth"] = r.get_image().copy(), r.get_depth().copy()
        np.savez(os.path.join(out_dir, "out.npz"), **results)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_strips_gather_and_peer_store(tmp_path):
    from minirender_b200 import cabi
    if cabi.load().mr_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    z = np.load(str(tmp_path / "out.npz"))
    assert (z["full_depth"] < 1e10).sum() > 10000
    for k in ("gather", "peer"):
        assert (z[k + "_depth"].view(np.uint32) == z["full_depth"].view(np.uint32)).all(), k
        assert (z[k + "_image"].view(np.uint32) == z["full_image"].view(np.uint32)).all(), k
