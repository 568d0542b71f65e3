"""GPU tests of the multi-GPU frame exchange.

The fused kernel (b200rt_finalize_peers_device) only sees device pointers, so one GPU is enough to
check its arithmetic: N "peer" sum buffers on the same device, every rank's launch issued in turn,
result compared bit for bit with "add in rank order, scale, b200rt_tonemap".  The real thing —
symmetric memory across processes, device barriers, NVLink — runs under torchrun when the box has
two or more GPUs (skipped otherwise) and must agree bit for bit with the NCCL reduce path at
world size 2 (a + b is commutative) and to FP32 summation order beyond.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_ranks(torch, capi, frames, n_px, spp, root_hdr, root_ldr, clamp=0):
    lib = capi.lib()
    ptrs = (C.c_void_p * len(frames))(*[f.data_ptr() for f in frames])
    for rank in range(len(frames)):
        capi._check(lib.b200rt_finalize_peers_device(ptrs, len(frames), rank, n_px, 1.0 / spp, root_hdr.data_ptr(),
                                                     root_ldr.data_ptr() if root_ldr is not None else None, clamp, 0, None))
    torch.cuda.synchronize()


@pytest.mark.parametrize("n_peers", [1, 2, 3, 8, 16])
@pytest.mark.parametrize("n_px", [1, 5, 1023, 1024, 1025, 4099, 64 * 36, 1920 * 1080 + 3])
def test_peer_finalize_matches_sum_scale_tonemap(n_peers, n_px):
    import torch
    from cpp_raytracer_b200 import capi
    g = torch.Generator(device="cuda").manual_seed(n_peers * 1000003 + n_px)
    frames = [torch.rand(n_px * 3, generator=g, device="cuda", dtype=torch.float32) * 40.0 for _ in range(n_peers)]
    for f in frames:
        f[:2] = 0.0
    frames[-1][2] = 1e4                                             # a saturated blue pixel: no clamp, value beyond 255
    spp = 37
    want = frames[0].clone()
    for f in frames[1:]:
        want += f                                                   # rank order, FP32, like the kernel
    want *= np.float32(1.0 / spp)
    hdr = torch.full((n_px * 3 + 8,), -7.0, device="cuda")          # separate root buffers, with guard words
    ldr = torch.full((n_px * 3 + 8,), -7, device="cuda", dtype=torch.int32)
    _run_ranks(torch, capi, frames, n_px, spp, hdr, ldr)
    assert torch.equal(hdr[: n_px * 3], want)
    assert bool((hdr[n_px * 3:] == -7.0).all()) and bool((ldr[n_px * 3:] == -7).all())
    want_ldr = capi.tonemap(want.cpu().numpy().reshape(-1, 3))
    assert np.array_equal(ldr[: n_px * 3].cpu().numpy().reshape(-1, 3), want_ldr)
    assert want_ldr.max() > 255                                     # the unclamped case was exercised

    # in place (root_hdr aliases peer 0's sum buffer, as FrameRenderer uses it), no tone map, clamp variant
    inplace = [f.clone() for f in frames]
    _run_ranks(torch, capi, inplace, n_px, spp, inplace[0], None)
    assert torch.equal(inplace[0], want)
    for a, b in zip(inplace[1:], frames[1:]):
        assert torch.equal(a, b)                                    # other ranks' buffers untouched
    _run_ranks(torch, capi, frames, n_px, spp, hdr, ldr, clamp=1)
    assert np.array_equal(ldr[: n_px * 3].cpu().numpy().reshape(-1, 3), np.clip(want_ldr, 0, 255))


def test_peer_finalize_rejects_bad_arguments():
    import torch
    from cpp_raytracer_b200 import capi
    lib = capi.lib()
    f = torch.zeros(64 * 3, device="cuda")
    ptrs = (C.c_void_p * 2)(f.data_ptr(), f.data_ptr())
    ok = lambda *a: lib.b200rt_finalize_peers_device(*a)
    assert ok(ptrs, 0, 0, 64, 1.0, f.data_ptr(), None, 0, 0, None) != 0          # no peers
    assert ok(ptrs, 17, 0, 64, 1.0, f.data_ptr(), None, 0, 0, None) != 0         # too many
    assert ok(ptrs, 2, 2, 64, 1.0, f.data_ptr(), None, 0, 0, None) != 0          # rank out of range
    assert ok(ptrs, 2, 0, 64, 1.0, None, None, 0, 0, None) != 0                  # no root
    mis = (C.c_void_p * 2)(f.data_ptr() + 4, f.data_ptr())
    assert ok(mis, 2, 0, 64, 1.0, f.data_ptr(), None, 0, 0, None) != 0           # misaligned peer
    assert b"16-byte" in lib.b200rt_last_error()
    assert ok(ptrs, 2, 0, 0, 1.0, None, None, 0, 0, None) == 0                   # empty frame is a no-op


def test_two_gpu_peer_exchange_matches_nccl_reduce(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = tmp_path / "exchange.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(REPO, "scripts", "exchange_check.py"), "--out", str(out)]
    r = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import json
    res = json.loads(out.read_text())
    assert res["hdr_equal"] and res["ldr_equal"], res
