"""Row-band split of one pair: host logic on CPU, the full schedule on one GPU with the
ranks emulated as handles (halos copied plane to plane), bit-exact against the single-GPU path."""
import numpy as np
import pytest

from ug_stereomatcher_b200 import rowband, synth


def test_band_rows_cover_every_level(oracle):
    w, h = oracle.level_dims(4928, 3264)
    for world in (1, 2, 4, 8):
        for lv in range(3):
            bands = [rowband.band_rows(h[lv], world, r) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == h[lv]
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            assert min(e - a for a, e in bands) >= rowband.HALO


def test_halo_is_large_enough():
    # per iteration: K1 on the band +- (S+1), reading state rows +- 3 more; S = 10 on the two finest levels
    assert rowband.HALO >= (10 + 1) + 3


@pytest.mark.parametrize("shape", [(3264, 4928), (1632, 2464), (816, 1232)])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_upsample_reads_only_valid_coarse_rows(oracle, shape, world):
    w, h = oracle.level_dims(shape[1], shape[0])
    for lv in (1, 0):  # fine level lv is upsampled from lv+1
        for r in range(world):
            ya_f, yb_f = rowband.band_rows(h[lv], world, r)
            ya_c, yb_c = rowband.band_rows(h[lv + 1], world, r)
            ok, _ = rowband._upsample_rows_ok(ya_f, yb_f, h[lv], ya_c, yb_c, h[lv + 1])
            assert ok, (shape, world, lv, r)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 4])
def test_emulated_rowband_bit_exact(capi, world):
    import torch

    rows, cols = 816, 1232
    L, R, _, _ = synth.make_pair(cols, rows, 4)
    dev = torch.device("cuda", 0)
    dL, dR = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
    out = [torch.empty((rows, cols), dtype=torch.float32, device=dev) for _ in range(3)]
    handles = [capi.Handle() for _ in range(world)]
    rowband.run_emulated(handles, dL.data_ptr(), dR.data_ptr(), rows, cols, cols * 3, out, split_levels=3)
    torch.cuda.synchronize()
    with capi.Handle() as h:
        want = h.match_full(L, R)
    for o, w_ in zip(out, want):
        got = o.cpu().numpy()
        assert (got == w_).all(), "row-band (world=%d) differs in %d px" % (world, int((got != w_).sum()))
    for hd in handles:
        hd.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_cpp_rowband_driver_bit_exact(capi, world):
    """ugsm_rowband_match (csrc/ugsm_rowband.cu): the C++ driver -- one host thread per rank, halo rows copied straight
    between the ranks' level planes, event-ordered streams -- with all ranks emulated on cuda:0 (same-device copies
    instead of NVLink peer copies; the multi-GPU run is tools/rowband_bench.py / bench.py --mode rowband).  Two pairs
    through the same handles (graph replay of the replicated part), bit-exact against the single-GPU path."""
    import torch

    rows, cols = 816, 1232
    dev = torch.device("cuda", 0)
    handles = [capi.Handle() for _ in range(world)]
    out = [torch.empty((rows, cols), dtype=torch.float32, device=dev) for _ in range(3)]
    for seed in (4, 5):
        L, R, _, _ = synth.make_pair(cols, rows, seed)
        dL, dR = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
        for o in out:
            o.fill_(float("nan"))
        torch.cuda.synchronize()
        ms = capi.rowband_match(handles, [0] * world, [dL.data_ptr()] * world, [dR.data_ptr()] * world, rows, cols,
                                cols * 3, [o.data_ptr() for o in out])
        assert ms > 0
        with capi.Handle() as h:
            want = h.match_full(L, R)
        for o, w_ in zip(out, want):
            got = o.cpu().numpy()
            assert (got == w_).all(), "C++ row-band (world=%d) differs in %d px" % (world, int((got != w_).sum()))
    for hd in handles:
        hd.close()
