"""world_size-2 `gloo` test of the N>1 path on CPU: interleaved-tile partition + the single reduce of the PixelData
accumulator + ray-counter sum.  The per-rank renderer here is the CPU oracle (the GPU renderer is exercised by
tests/test_gpu_parity.py::test_window_and_tiles_cover_image_exactly and by `bench.py --gpus N`)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, TW, TH, SPP, DEPTH = 80, 48, 16, 16, 2, 6


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cudatracerlib_b200 as ctl
    from cudatracerlib_b200 import api
    import oracle_binding as ob
    scene = ctl.Scene("cornell7", W, H)
    accum = torch.zeros(H * W * 7, dtype=torch.float32)
    img = accum.numpy().view(api.PIXEL_DTYPE).reshape(H, W)
    rays = [0]

    def render_pass(p, new_trace):
        if new_trace:
            accum.zero_(); rays[0] = 0
        for win in ctl.tiles_of_rank(W, H, TW, TH, rank, world):
            _, r = ob.render(scene.view, W, H, n_passes=1, pass_first=p, max_path_length=DEPTH, window=win, n_threads=1, img=img)
            rays[0] += r

    df = ctl.DistributedFrame(accum, render_pass, lambda: rays[0])
    df.frame(SPP)
    total = df.total_rays()
    # ownership: before the reduce each rank only touched its own tiles (+ the 1-pixel jitter fringe)
    if rank == 0:
        np.save(out_path, accum.numpy())
        with open(out_path + ".rays", "w") as f:
            f.write(str(total))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tiled_render_equals_single_rank(tmp_path, orc):
    import cudatracerlib_b200 as ctl
    from cudatracerlib_b200 import api
    out = str(tmp_path / "accum.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out).view(api.PIXEL_DTYPE).reshape(H, W)
    scene = ctl.Scene("cornell7", W, H)
    ref, ref_rays = orc.render(scene.view, W, H, n_passes=SPP, max_path_length=DEPTH, n_threads=2)
    assert int(open(out + ".rays").read()) == ref_rays               # ray counters sum to the single-rank count
    assert np.array_equal(got["weight_sum"], ref["weight_sum"])       # every path landed exactly once
    assert got["weight_sum"].sum() == SPP * W * H
    assert np.allclose(got["rgb"], ref["rgb"], rtol=1e-6, atol=1e-7)  # adding the other rank's zeros is exact; spill pixels may reorder
    exact = (got["rgb"] == ref["rgb"]).all(axis=2).mean()
    assert exact > 0.98


def test_tiles_of_rank_matches_tile_owner():
    import cudatracerlib_b200 as ctl
    for (w, h, tw, th, n) in ((80, 48, 16, 16, 2), (100, 70, 16, 16, 3), (1920, 1080, 64, 64, 8)):
        own = ctl.tile_owner(w, h, tw, th, n)
        cover = np.full((h, w), -1)
        for r in range(n):
            for (x0, y0, x1, y1) in ctl.tiles_of_rank(w, h, tw, th, r, n):
                assert (cover[y0:y1, x0:x1] == -1).all()
                cover[y0:y1, x0:x1] = r
        assert np.array_equal(cover, own)
        counts = np.bincount(own.ravel(), minlength=n)
        assert counts.min() > 0.8 * counts.max() or n > 4   # interleaving balances the load


def _worker_passes(rank, world, port, out_path, spp):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cudatracerlib_b200 as ctl
    from cudatracerlib_b200 import api
    import oracle_binding as ob
    scene = ctl.Scene("cornell7", W, H)
    accum = torch.zeros(H * W * 7, dtype=torch.float32)
    img = accum.numpy().view(api.PIXEL_DTYPE).reshape(H, W)
    rays = [0]

    def render_pass(p, new_trace):   # the oracle's WavefrontPathTracer renders pass p of the frame
        if new_trace:
            accum.zero_(); rays[0] = 0
        _, r, _ = ob.render_wavefront(scene.view, W, H, n_passes=1, pass_first=p, max_path_length=DEPTH, img=img)
        rays[0] += r

    dp = ctl.DistributedPasses(accum, render_pass, lambda: rays[0])
    assert dp.passes_of_rank(spp) == list(range(rank, spp, world))
    dp.frame(spp)
    total = dp.total_rays()
    if rank == 0:
        np.save(out_path, accum.numpy())
        with open(out_path + ".rays", "w") as f:
            f.write(str(total))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("spp", [3, 1])
def test_two_rank_pass_sharing_equals_single_rank(tmp_path, orc, spp):
    """WavefrontPathTracer frames shared by pass index (DistributedPasses): ranks render passes 0, 2 / 1 (or, with one pass, rank 1 adds
    zeros); the reduced image equals the single-rank frame up to the order of the per-pixel additions, ray counters add up."""
    import cudatracerlib_b200 as ctl
    from cudatracerlib_b200 import api
    out = str(tmp_path / "accum_p.npy")
    mp.spawn(_worker_passes, args=(2, _free_port(), out, spp), nprocs=2, join=True)
    got = np.load(out).view(api.PIXEL_DTYPE).reshape(H, W)
    scene = ctl.Scene("cornell7", W, H)
    ref, ref_rays, _ = orc.render_wavefront(scene.view, W, H, n_passes=spp, max_path_length=DEPTH)
    assert int(open(out + ".rays").read()) == ref_rays
    assert np.array_equal(got["weight_sum"], ref["weight_sum"]) and got["weight_sum"].sum() == spp * W * H
    assert np.allclose(got["rgb"], ref["rgb"], rtol=1e-6, atol=1e-7)
