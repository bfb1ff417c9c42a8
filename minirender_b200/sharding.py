"""How the render path shards across the GPUs of one box. Pure host logic plus
torch.distributed plumbing; no pixel is computed here.

  view batch (BASELINE.json configs[4])  every rank holds a replica of the scene and renders its own
      views; nothing is exchanged on the data path.
  strips (configs[2])  one frame, rank r renders image rows strip_rows(h, r, N); rows of the edge
      chain are independent and winners are order-independent given submission ids, so the union
      of the strips is byte-identical to the single-GPU frame. Finished strips reach rank 0 either
      by NCCL send/recv (gather_strips) or, fused with the tile store, directly through peer memory
      over NVLink (open_peer_target + Renderer row range: the rasterizer's stores land in rank 0's
      framebuffer, no separate gather step).
"""
import ctypes as C

import numpy as np

TILE = 16


def views_for_rank(n_views, rank, world):
    """Interleaved assignment: rank r renders views r, r+N, r+2N, ..."""
    return list(range(rank, n_views, world))


def strip_rows(height, rank, world, tile=TILE):
    """Contiguous horizontal strips of whole tile rows, as even as possible. Returns (begin, end);
    begin == end for ranks left without rows (more ranks than tile rows)."""
    tile_rows = (height + tile - 1) // tile
    base, extra = divmod(tile_rows, world)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return min(first * tile, height), min((first + count) * tile, height)


def all_strips(height, world, tile=TILE):
    return [strip_rows(height, r, world, tile) for r in range(world)]


def merge_row_ranges(ranges):
    """Adjacent or overlapping (begin, end) row ranges joined: the peers of a gathering rank usually own one contiguous
    block of rows, which is then one clear instead of one per peer."""
    out = []
    for b, e in sorted((b, e) for b, e in ranges if e > b):
        if out and b <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], e))
        else:
            out.append((b, e))
    return out


def balanced_strips(height, strips, times, tile=TILE):
    """Re-cuts `strips` (a list of (begin, end) covering [0, height) in whole tile rows) so that every rank's strip
    takes about the same time, given the time each rank measured for its current strip: the cost of a tile row is
    taken as constant within a strip (time / tile rows), and the new cuts split the accumulated cost evenly. Every
    rank keeps at least one tile row. Pure host logic: every rank calls it with the same (all-gathered) times and
    gets the same answer; a few rounds converge, also when part of a rank's time does not depend on its rows."""
    tile_rows = (height + tile - 1) // tile
    world = len(strips)
    if world <= 1 or tile_rows < world:
        return list(strips)
    cost = [0.0] * tile_rows
    for (b, e), t in zip(strips, times):
        n = (e - b + tile - 1) // tile
        for k in range(n):
            cost[b // tile + k] = max(float(t), 1e-9) / n
    total = sum(cost)
    cuts, acc, row = [0], 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while row < tile_rows and acc + cost[row] * 0.5 < target:
            acc += cost[row]
            row += 1
        row = max(row, cuts[-1] + 1)                 # at least one tile row per rank ...
        row = min(row, tile_rows - (world - r))      # ... for the ranks still to come as well
        acc = sum(cost[:row])
        cuts.append(row)
    cuts.append(tile_rows)
    return [(min(cuts[r] * tile, height), min(cuts[r + 1] * tile, height)) for r in range(world)]


class DeviceArray:
    """Wraps a raw device pointer for torch (``torch.as_tensor(DeviceArray(...), device='cuda')``)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 2}


def device_tensors(ctx_lib, ctx, h, w, device):
    """The context's float image (h,w,3) and depth (h,w) buffers as torch tensors (zero copy)."""
    import torch
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert ctx_lib.mr_device_buffers(ctx, C.byref(a), C.byref(b), C.byref(c)) == 0
    img = torch.as_tensor(DeviceArray(a.value, (h, w, 3)), device=device)
    dep = torch.as_tensor(DeviceArray(b.value, (h, w)), device=device)
    return img, dep


def gather_strips(image, depth, height, rank, world, dist, dst=0):
    """Moves every rank's finished rows into rank `dst`'s framebuffer with point-to-point
    send/recv (NCCL on GPUs, gloo on CPU tensors). `image` (h,w,3) and `depth` (h,w) are each rank's
    full-size buffers; only the rank's own rows are valid on input. Works on torch tensors."""
    strips = all_strips(height, world)
    ops = []
    if rank == dst:
        for r, (rb, re) in enumerate(strips):
            if r == dst or re <= rb:
                continue
            ops.append(dist.P2POp(dist.irecv, image[rb:re], r))
            ops.append(dist.P2POp(dist.irecv, depth[rb:re], r))
    else:
        rb, re = strips[rank]
        if re > rb:
            ops.append(dist.P2POp(dist.isend, image[rb:re], dst))
            ops.append(dist.P2POp(dist.isend, depth[rb:re], dst))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def open_peer_target(ctx_lib, ctx, rank, world, dist, dst=0):
    """Makes this rank's tile stores land in rank `dst`'s framebuffer (peer memory over NVLink).
    Returns a closer callable. Needs one process per GPU on one node."""
    handles = (C.c_ubyte * 128)()
    if rank == dst:
        assert ctx_lib.mr_ipc_export(ctx, handles) == 0
    payload = [bytes(handles) if rank == dst else None]
    dist.broadcast_object_list(payload, src=dst)
    if rank == dst:
        return lambda: None
    buf = (C.c_ubyte * 128).from_buffer_copy(payload[0])
    pi, pd = C.c_void_p(), C.c_void_p()
    rc = ctx_lib.mr_ipc_open(ctx, buf, C.byref(pi), C.byref(pd))
    if rc != 0:
        raise RuntimeError("mr_ipc_open failed: %s" % ctx_lib.mr_last_error(ctx).decode())
    assert ctx_lib.mr_set_remote_target(ctx, pi, pd) == 0

    def close():
        ctx_lib.mr_set_remote_target(ctx, None, None)
        ctx_lib.mr_ipc_close(ctx, pi, pd)
    return close


def open_peer_targets(ctx_lib, ctx, rank, world, dist, dst=0):
    """Two-framebuffer form of open_peer_target: rank `dst` (with two output slots) exports both of its framebuffers;
    returns (targets, first, close): targets[s] = (d_image, d_depth) of dst's slot s as seen from this rank (None on dst),
    first = the slot dst's next frame will be rendered into (its frames then alternate)."""
    payload = [None]
    if rank == dst:
        hs = []
        for slot in (0, 1):
            h = (C.c_ubyte * 128)()
            assert ctx_lib.mr_ipc_export_slot(ctx, slot, h) == 0, ctx_lib.mr_last_error(ctx)
            hs.append(bytes(h))
        payload = [(hs, (ctx_lib.mr_output_slot(ctx) + 1) % 2)]
    dist.broadcast_object_list(payload, src=dst)
    hs, first = payload[0]
    if rank == dst:
        return None, first, (lambda: None)
    targets = []
    for h in hs:
        buf = (C.c_ubyte * 128).from_buffer_copy(h)
        pi, pd = C.c_void_p(), C.c_void_p()
        if ctx_lib.mr_ipc_open(ctx, buf, C.byref(pi), C.byref(pd)) != 0:
            raise RuntimeError("mr_ipc_open failed: %s" % ctx_lib.mr_last_error(ctx).decode())
        targets.append((pi, pd))

    def close():
        ctx_lib.mr_set_remote_target(ctx, None, None)
        for pi, pd in targets:
            ctx_lib.mr_ipc_close(ctx, pi, pd)
    return targets, first, close


class StripJoin:
    """Device-side join of a strip-sharded frame (mr_stream_signal / mr_stream_wait): no collective and no host
    synchronisation on the per-frame path. Rank `dst` owns the words: arrived[r] (written by rank r over NVLink after
    its rows have landed in dst's framebuffer) and go (written by dst when a frame is complete and may be overwritten).

        join = StripJoin(lib, ctx, rank, world, dist)
        for i in range(frames):
            join.begin(i)        # peers: wait until dst has released the framebuffer of frame i - 1
            renderer.render()    # rows of this rank, stored straight into dst's framebuffer
            join.end(i)          # signal arrival; dst: wait for every rank, (consume the frame,) release
    """

    def __init__(self, ctx_lib, ctx, rank, world, dist, dst=0, targets=None, first=None):
        """targets / first (from open_peer_targets): dst alternates between two framebuffers, frame i goes to slot
        (first + i) % 2. The peers may then store frame i + 1 while dst still waits for, consumes and clears frame i:
        a peer's tile kernel for frame i only waits until frame i - 2 (the previous user of that buffer) is released."""
        self.lib, self.ctx, self.rank, self.world, self.dst = ctx_lib, ctx, rank, world, dst
        self.targets, self.first, self.double = targets, first, first is not None
        self.words = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        if rank == dst:
            assert ctx_lib.mr_sync_words(ctx, world + 1, C.byref(self.words)) == 0
            assert ctx_lib.mr_ipc_export_ptr(ctx, self.words, handle) == 0
        payload = [bytes(handle) if rank == dst else None]
        dist.broadcast_object_list(payload, src=dst)
        self.opened = False
        if rank != dst:
            buf = (C.c_ubyte * 64).from_buffer_copy(payload[0])
            rc = ctx_lib.mr_ipc_open_ptr(ctx, buf, C.byref(self.words))
            if rc != 0:
                raise RuntimeError("mr_ipc_open_ptr failed: %s" % ctx_lib.mr_last_error(ctx).decode())
            self.opened = True

    def _word(self, i):
        return C.c_void_p(self.words.value + 4 * i)

    def begin(self, frame):
        # a peer's geometry may run ahead; only its tile kernel (which stores into dst's framebuffer) waits for dst
        if self.rank == self.dst:
            return
        if self.double:
            pi, pd = self.targets[(self.first + frame) % 2]
            assert self.lib.mr_set_remote_target(self.ctx, pi, pd) == 0
            if frame > 1:
                assert self.lib.mr_set_raster_gate(self.ctx, self._word(self.world), frame - 1) == 0
        elif frame > 0:
            assert self.lib.mr_set_raster_gate(self.ctx, self._word(self.world), frame) == 0

    def end(self, frame, clear_rows=None):
        """clear_rows = (background rgb, [(row_begin, row_end), ...]) on dst: with sparse remote stores the peers only
        write the tiles they drew into, so dst resets their rows to the clear values before it lets them in again."""
        if self.rank != self.dst:
            assert self.lib.mr_stream_signal(self.ctx, self._word(self.rank), frame + 1) == 0
            return
        # dst's own rows are ordered by its stream: it waits for the others' words only (the words before and behind
        # its own) and clears the peers' rows in as few launches as they form blocks: the gathering rank enqueues
        # the most per frame, and its host time per frame is part of the frame rate at eight ranks
        if self.dst > 0:
            assert self.lib.mr_stream_wait(self.ctx, self._word(0), self.dst, frame + 1) == 0
        if self.dst + 1 < self.world:
            assert self.lib.mr_stream_wait(self.ctx, self._word(self.dst + 1), self.world - self.dst - 1, frame + 1) == 0
        # (a consumer of the assembled frame would be enqueued here)
        if clear_rows is not None:
            bg = (C.c_float * 3)(*clear_rows[0])
            for rb, re in merge_row_ranges(clear_rows[1]):
                if self.double:
                    assert self.lib.mr_clear_rows_slot(self.ctx, (self.first + frame) % 2, bg, rb, re) == 0
                else:
                    assert self.lib.mr_clear_rows(self.ctx, bg, rb, re) == 0
        assert self.lib.mr_stream_signal(self.ctx, self._word(self.world), frame + 1) == 0

    def close(self):
        if self.opened:
            self.lib.mr_ipc_close_ptr(self.ctx, self.words)
            self.opened = False
