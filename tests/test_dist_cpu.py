"""CPU, world_size 2 over gloo: the multi-GPU host logic (strip partition + gather into rank 0,
interleaved view batch) with the CPU oracle standing in for the per-rank renderer."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import minirender_b200 as m
    from minirender_b200 import scenes, sharding
    import pyoracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    be = m.Backend()
    setup = scenes.SMALL_SCENES["clip"](be)
    r = setup.apply(m.Renderer(be))
    h, w = setup.height, setup.width
    # --- strips: every rank renders only its rows, rank 0 ends up with the whole frame ---
    rb, re = sharding.strip_rows(h, rank, world)
    r.set_row_range(rb, re)
    r.prepare()
    bufs = dict(image=np.full((h, w, 3), -7.0, np.float32), depth=np.full((h, w), -7.0, np.float32))
    pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), w, h, into=bufs)
    assert (bufs["depth"][:rb] == -7.0).all() and (bufs["depth"][re:] == -7.0).all()  # other rows untouched
    img, dep = torch.from_numpy(bufs["image"]), torch.from_numpy(bufs["depth"])
    sharding.gather_strips(img, dep, h, rank, world, dist, dst=0)
    # --- view batch: interleaved views, per-rank checksums reduced on rank 0 ---
    r.set_row_range(0, 0)
    sums = torch.zeros(6, dtype=torch.float64)
    for v in sharding.views_for_rank(6, rank, world):
        r.set_view(be.mul(be.translate(0, 0, -20), be.rotate_x(np.float32(-0.3)), be.rotate_z(np.float32(0.2 * v))))
        r.prepare()
        o = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), w, h)
        sums[v] = float(o["depth"][o["depth"] < 1e10].astype(np.float64).sum())
    dist.all_reduce(sums)
    if rank == 0:
        np.savez(os.path.join(out_dir, "rank0.npz"), image=img.numpy(), depth=dep.numpy(), sums=sums.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_strips_and_views(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(str(tmp_path / "rank0.npz"))
    sys.path[:0] = [os.path.join(ROOT, "oracle")]
    import minirender_b200 as m
    from minirender_b200 import scenes
    import pyoracle
    be = m.Backend()
    setup = scenes.SMALL_SCENES["clip"](be)
    r = setup.apply(m.Renderer(be))
    r.prepare()
    full = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height)
    assert (z["depth"].view(np.uint32) == full["depth"].view(np.uint32)).all()
    assert (z["image"].view(np.uint32) == full["image"].view(np.uint32)).all()
    for v in range(6):
        r.set_view(be.mul(be.translate(0, 0, -20), be.rotate_x(np.float32(-0.3)), be.rotate_z(np.float32(0.2 * v))))
        r.prepare()
        o = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height)
        assert z["sums"][v] == float(o["depth"][o["depth"] < 1e10].astype(np.float64).sum())
