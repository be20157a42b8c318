"""Row-band split of ONE pair across the GPUs of a box (second way the path shards).

Levels >= `split_levels` (small: 32.7 % of the pixel-iterations at 16 MP for split_levels = 3) are
computed replicated on every rank; the finest levels are split into contiguous row bands, one per
rank.  Per iteration a rank runs K1 on its band extended by S+1 rows (so K2 needs no exchange of
K1's output), K2 on its band, and then exchanges HALO rows of the new (H, V, C) with its two
neighbours -- the only communication of the data path.  Disparity is horizontal-dominant, which is
why the bands are rows.  Results are bit-identical to the single-GPU path (tests/test_rowband.py).

Two transports drive the same schedule:
  * run_distributed(): one process per GPU, halo exchange with torch.distributed P2P (NCCL over
    NVLink; gloo is not applicable, the planes live in device memory);
  * run_emulated():    all "ranks" are handles on ONE GPU and halos are copied between their planes
    directly -- the way the schedule is verified on a single GPU.
"""
import numpy as np

HALO = 16  # rows exchanged per side and iteration: >= (S+1) + 3 for S = 10, with slack for the upsample


def band_rows(height, world, rank):
    return (height * rank) // world, (height * (rank + 1)) // world


def _upsample_rows_ok(ya_f, yb_f, h_f, ya_c, yb_c, h_c):
    """The fine rows [ya_f-HALO, yb_f+HALO) must only read coarse rows that are valid on this rank."""
    sf = np.float32(1 / 1.41421356)
    lo, hi = max(0, ya_f - HALO), min(h_f, yb_f + HALO)
    src_lo = int(np.floor((np.float32(lo) + np.float32(0.5)) * sf))
    src_hi = min(h_c - 1, int(np.floor((np.float32(hi - 1) + np.float32(0.5)) * sf)))
    return src_lo >= max(0, ya_c - HALO) and src_hi <= min(h_c, yb_c + HALO) - 1, (lo, hi)


class RankState:
    """Everything one rank needs: its handle and the row bands of the split levels."""

    def __init__(self, handle, rank, world, split_levels):
        self.h, self.rank, self.world, self.split = handle, rank, world, split_levels

    def begin(self, d_rgbL, d_rgbR, rows, cols, step):
        self.h.band_begin(d_rgbL, d_rgbR, rows, cols, step, self.split)

    def level(self, level):
        b = self.h.band_level_info(level)
        ya, yb = band_rows(b.height, self.world, self.rank)
        if self.world > 1 and yb - ya < HALO:
            raise ValueError("band of %d rows at level %d is smaller than the halo" % (yb - ya, level))
        return b, level, ya, yb


def _schedule(states, exchange, gather):
    """The shared schedule.  `states`: the ranks driven by this process (1 when distributed)."""
    split = states[0].split
    prev = {}
    for level in range(split - 1, -1, -1):
        cur = {}
        for s in states:
            b, lv, ya, yb = s.level(level)
            cur[s.rank] = (b, lv, ya, yb)
            if level < split - 1:
                pb, _, pya, pyb = prev[s.rank]
                ok, (lo, hi) = _upsample_rows_ok(ya, yb, b.height, pya, pyb, pb.height)
                if not ok:
                    raise AssertionError("upsample would read coarse rows outside the valid halo")
                s.h.band_upsample(level + 1, lo, hi)
        b0 = cur[states[0].rank][0]
        for m in range(1, b0.iters + 1):
            for s in states:
                b, lv, ya, yb = cur[s.rank]
                s.h.band_iteration(level, m, ya, yb, b.smooth + 1)
            exchange(cur)
        prev = cur
    gather(prev)


def run_emulated(handles, d_rgbL, d_rgbR, rows, cols, step, out, split_levels=3):
    """All ranks on one GPU (handles[r] = rank r); halos copied plane to plane.  `out`: 3 dense
    device tensors (rows x cols) that receive rank 0's assembled H, V, C."""
    import torch

    device = out[0].device
    world = len(handles)
    states = [RankState(h, r, world, split_levels) for r, h in enumerate(handles)]
    # one explicit stream orders the work of all emulated ranks (stream handle 0 would mean "the
    # handle's own stream" to ugsm_set_stream, i.e. no ordering between handles)
    side = torch.cuda.Stream(device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        _run_emulated_on(side, states, handles, d_rgbL, d_rgbR, rows, cols, step, out, device, world)
    torch.cuda.current_stream(device).wait_stream(side)
    for h in handles:
        h.set_stream(0)


def _run_emulated_on(side, states, handles, d_rgbL, d_rgbR, rows, cols, step, out, device, world):
    import torch

    for s in states:
        s.h.set_stream(side.cuda_stream)
        s.begin(d_rgbL, d_rgbR, rows, cols, step)

    def stage(b, nrows):
        return torch.empty((3, nrows, b.pitch), dtype=torch.float32, device=device)

    def exchange(cur):
        for r in range(world - 1):  # neighbours r (upper band) and r+1 (lower band)
            b, lv, _, yb = cur[r]
            _, _, ya, _ = cur[r + 1]
            assert ya == yb
            buf = stage(b, HALO)
            handles[r].band_pack(lv, ya - HALO, ya, buf.data_ptr())      # bottom rows of r -> top halo of r+1
            handles[r + 1].band_unpack(lv, ya - HALO, ya, buf.data_ptr())
            buf2 = stage(b, HALO)
            handles[r + 1].band_pack(lv, ya, ya + HALO, buf2.data_ptr())  # top rows of r+1 -> bottom halo of r
            handles[r].band_unpack(lv, ya, ya + HALO, buf2.data_ptr())

    def gather(cur):
        for r in range(1, world):
            b, lv, ya, yb = cur[r]
            buf = stage(b, yb - ya)
            handles[r].band_pack(lv, ya, yb, buf.data_ptr())
            handles[0].band_unpack(lv, ya, yb, buf.data_ptr())
        states[0].h.band_emit(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), 0, rows)

    _schedule(states, exchange, gather)


def run_distributed(handle, d_rgbL, d_rgbR, rows, cols, step, out, split_levels=3):
    """One process per GPU under torchrun (NCCL).  Rank 0's `out` receives the assembled planes."""
    import torch
    import torch.distributed as dist

    device = out[0].device
    # everything (kernels, packing copies, NCCL P2P) is ordered on one explicit side stream
    side = torch.cuda.Stream(device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        _run_distributed_on(side, handle, d_rgbL, d_rgbR, rows, cols, step, out, split_levels, device)
    torch.cuda.current_stream(device).wait_stream(side)
    handle.set_stream(0)


def _run_distributed_on(side, handle, d_rgbL, d_rgbR, rows, cols, step, out, split_levels, device):
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    handle.set_stream(side.cuda_stream)
    st = RankState(handle, rank, world, split_levels)
    st.begin(d_rgbL, d_rgbR, rows, cols, step)

    def stage(b, nrows):
        return torch.empty((3, nrows, b.pitch), dtype=torch.float32, device=device)

    def exchange(cur):
        b, lv, ya, yb = cur[rank]
        ops, recvs = [], []
        if rank > 0:
            send, recv = stage(b, HALO), stage(b, HALO)
            handle.band_pack(lv, ya, ya + HALO, send.data_ptr())
            ops += [dist.P2POp(dist.isend, send, rank - 1), dist.P2POp(dist.irecv, recv, rank - 1)]
            recvs.append((recv, ya - HALO, ya))
        if rank < world - 1:
            send, recv = stage(b, HALO), stage(b, HALO)
            handle.band_pack(lv, yb - HALO, yb, send.data_ptr())
            ops += [dist.P2POp(dist.isend, send, rank + 1), dist.P2POp(dist.irecv, recv, rank + 1)]
            recvs.append((recv, yb, yb + HALO))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for recv, lo, hi in recvs:
            handle.band_unpack(lv, lo, hi, recv.data_ptr())

    def gather(cur):
        b, lv, ya, yb = cur[rank]
        if rank == 0:
            for r in range(1, world):
                a, e = band_rows(b.height, world, r)
                buf = stage(b, e - a)
                dist.recv(buf, r)
                handle.band_unpack(lv, a, e, buf.data_ptr())
            handle.band_emit(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), 0, rows)
        else:
            buf = stage(b, yb - ya)
            handle.band_pack(lv, ya, yb, buf.data_ptr())
            dist.send(buf, 0)

    _schedule([st], exchange, gather)
