"""Multi-GPU rendering: sample split + sum-reduce (one process per GPU).

Every pixel-sample is independent (reference camera.h:286-289 loops over them serially), so
the frame shards by SAMPLES: rank k of G renders samples [k*spp/G, (k+1)*spp/G) of EVERY pixel
into a full-frame FP32 sum buffer on its own GPU (the scene + BVH are replicated; even the
3.1 M-sphere scene is ~150 MB).  The RNG is keyed by the global sample index, so the image does
not depend on G beyond FP32 summation order.  The only exchange step is the sum of the
W x H x 3 float frames onto rank 0, followed by /= spp and the tone map.  Two implementations:

  "peer"  every rank's sum buffer lives in symmetric (peer-mapped) memory; after a device barrier each
          rank runs libb200rt's reduce_finalize_peers kernel on ITS slice of the frame: it reads that
          slice from all ranks over NVLink, adds in rank order, scales, tone-maps and writes the result
          straight into rank 0's buffers (reduce-scatter + epilogue + gather as one launch per rank).
          Deterministic: the sum order is fixed by rank, not by arrival.
  "nccl"  dist.reduce to rank 0 (NCCL on GPUs; gloo in the CPU tests), then b200rt_finalize_device there.

PyTorch is plumbing here (device buffers, symmetric-memory rendezvous, streams, torch.distributed);
all rendering and the exchange kernel go through libb200rt.so.
"""
from __future__ import annotations

import numpy as np


def sample_range(spp: int, rank: int, world: int) -> tuple[int, int]:
    """(first sample, count) of `rank`; ranges tile [0, spp) exactly for any world size."""
    lo = (spp * rank) // world
    hi = (spp * (rank + 1)) // world
    return lo, hi - lo


def reduce_frames(local_sum, dst: int = 0, group=None):
    """Sum-reduces the per-rank frame sums onto rank `dst` (in place).  `local_sum` is a torch
    tensor on the rank's device (cuda under NCCL, cpu under gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(local_sum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return local_sum


class PeerFrames:
    """Symmetric-memory frame buffers of one rank: [H*W*3 float32 sums | H*W*3 int32 tone-mapped], mapped
    into every peer's address space (torch.distributed._symmetric_memory: cuMem + NVLink P2P)."""

    def __init__(self, h: int, w: int, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        n = h * w * 3
        n_pad = (n + 3) // 4 * 4                      # keeps the int32 half 16-byte aligned
        self.group = group if group is not None else dist.group.WORLD
        self.buf = symm.empty(2 * n_pad, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.frame = self.buf[:n].view(h, w, 3)
        self.ldr = self.buf[n_pad:n_pad + n].view(torch.int32).view(h, w, 3)
        self.ldr_offset_bytes = n_pad * 4
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]   # rank r's buffer as addressable from this device
        self.world = len(self.ptrs)

    def barrier(self, channel: int = 0):
        """Device-side barrier across the ranks on the current stream (signal pads; no host sync)."""
        self.hdl.barrier(channel=channel)


class FrameRenderer:
    """Renders one frame of `scene` on this rank's GPU as its share of a world of ranks.

    render_sum()   -> enqueue this rank's samples into the device sum buffer (no sync)
    finish()       -> reduce to rank 0, scale by 1/spp; rank 0 returns the HDR frame tensor
    """

    def __init__(self, dev_scene, cam: np.ndarray, rank: int = 0, world: int = 1, device_index: int = 0,
                 seed: int = 0xB200, variant: int = 0, frame=None, ldr=None, peers: PeerFrames | None = None):
        """`frame` / `ldr`: optional preallocated device tensors (H x W x 3 float32 / int32) to reuse
        across frames instead of allocating per renderer.  `peers`: symmetric buffers (they replace
        frame / ldr) — finish() then uses the fused peer-memory exchange instead of dist.reduce."""
        import torch
        self.torch = torch
        self.scene = dev_scene
        self.cam = cam
        self.rank, self.world = rank, world
        self.seed, self.variant = seed, variant
        self.h, self.w = int(cam["image_h"][0]), int(cam["image_w"][0])
        self.spp = int(cam["spp"][0])
        self.first, self.count = sample_range(self.spp, rank, world)
        self.device = torch.device("cuda", device_index)
        self.peers = peers
        if peers is not None:
            frame, ldr = peers.frame, peers.ldr
        self.frame = frame if frame is not None else torch.empty((self.h, self.w, 3), dtype=torch.float32, device=self.device)
        self.ldr = ldr if ldr is not None else torch.empty((self.h, self.w, 3), dtype=torch.int32, device=self.device)

    def render_sum(self, want_stats: bool = False):
        """This rank's samples as a SUM frame.  FLAG_EXACT_COUNT: a rank whose share is empty (spp < world) renders
        nothing and gets a zero frame -- without it sample_count = 0 would mean "all of camera.spp"."""
        from . import capi
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        return self.scene.render_device(self.cam, self.frame.data_ptr(), stream=stream, seed=self.seed,
                                        sample_offset=self.first, sample_count=self.count, variant=self.variant,
                                        flags=capi.FLAG_SUM | capi.FLAG_EXACT_COUNT, want_stats=want_stats)

    def finish(self, tonemap: bool = True):
        from . import capi
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        if self.peers is not None and self.world > 1:
            import ctypes as C
            pf = self.peers
            ptrs = (C.c_void_p * pf.world)(*pf.ptrs)
            pf.barrier(0)                              # every rank's samples are in its sum buffer
            capi._check(capi.lib().b200rt_finalize_peers_device(
                ptrs, pf.world, self.rank, self.h * self.w, 1.0 / self.spp, pf.ptrs[0],
                pf.ptrs[0] + pf.ldr_offset_bytes if tonemap else None, 0, self.device.index, stream))
            pf.barrier(1)                              # every slice has landed in rank 0's buffers
            return self.frame if self.rank == 0 else None
        reduce_frames(self.frame, dst=0)
        if self.rank != 0:
            return None
        capi._check(capi.lib().b200rt_finalize_device(self.frame.data_ptr(), self.h * self.w, 1.0 / self.spp,
                                                      self.ldr.data_ptr() if tonemap else None, 0,
                                                      self.device.index, stream))
        return self.frame
