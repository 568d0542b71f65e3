"""world_size-2 `gloo` test of the multi-GPU path's host logic (no GPU): the sample split tiles
[0, spp), and sum-reducing the per-rank frame sums onto rank 0 gives the single-rank sum.  The
per-rank 'render' is stood in by a deterministic function of (pixel, sample) -- the real kernel
has the same property (streams keyed by the global sample index), which the GPU test
test_render_is_deterministic_and_split_invariant checks on the device."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_sample(pix, s):
    x = (pix * 2654435761 + s * 40503 + 12345) % 1000003
    return (x.astype(np.float64) / 1000003.0).astype(np.float32)


def _frame_sum(h, w, first, count):
    pix = np.arange(h * w, dtype=np.int64)
    acc = np.zeros(h * w, np.float32)
    for s in range(first, first + count):
        acc += _fake_sample(pix, s)
    return np.repeat(acc.reshape(h, w, 1), 3, axis=2)


def _worker(rank, world, port, spp, h, w, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpp_raytracer_b200.dist import reduce_frames, sample_range
    first, count = sample_range(spp, rank, world)
    local = torch.from_numpy(_frame_sum(h, w, first, count))
    reduce_frames(local, dst=0)
    if rank == 0:
        np.save(out_path, (local / spp).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sample_split_reduce_world2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    spp, h, w = 37, 6, 8      # odd spp: ranks get 18 and 19 samples
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(2, port, spp, h, w, out), nprocs=2, join=True)
    got = np.load(out)
    want = _frame_sum(h, w, 0, spp) / spp
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_sample_split_with_fewer_samples_than_ranks():
    """spp < world gives some ranks an EMPTY share; the shares must still tile [0, spp) exactly
    (FrameRenderer passes B200RT_FLAG_EXACT_COUNT so that count == 0 renders nothing instead of meaning "camera.spp";
    the device side of that is tests/test_gpu_multi.py::test_exact_count_zero_renders_nothing)."""
    from cpp_raytracer_b200 import capi
    from cpp_raytracer_b200.dist import sample_range
    for spp, world in [(4, 8), (1, 8), (0, 2), (3, 2), (7, 4), (1024, 8), (37, 2)]:
        covered = []
        for r in range(world):
            lo, n = sample_range(spp, r, world)
            assert n >= 0
            covered += list(range(lo, lo + n))
        assert covered == list(range(spp)), (spp, world)
    assert capi.FLAG_EXACT_COUNT == 8
    import inspect
    from cpp_raytracer_b200 import dist as rtdist
    assert "FLAG_EXACT_COUNT" in inspect.getsource(rtdist.FrameRenderer.render_sum)
