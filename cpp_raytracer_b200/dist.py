"""Multi-GPU rendering: sample split + sum-reduce (one process per GPU).

Every pixel-sample is independent (reference camera.h:286-289 loops over them serially), so
the frame shards by SAMPLES: rank k of G renders samples [k*spp/G, (k+1)*spp/G) of EVERY pixel
into a full-frame FP32 sum buffer on its own GPU (the scene + BVH are replicated; even the
3.1 M-sphere scene is ~150 MB).  The RNG is keyed by the global sample index, so the image does
not depend on G beyond FP32 summation order.  The only exchange step is one sum-reduce of the
W x H x 3 float frame to rank 0 (NCCL over NVLink under torchrun; gloo in the CPU tests), after
which rank 0 scales by 1/spp and tone-maps.  PyTorch is plumbing here (device buffers, streams,
torch.distributed); all rendering goes through libb200rt.so.
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


class FrameRenderer:
    """Renders one frame of `scene` on this rank's GPU as its share of a world of ranks.

    render_sum()   -> enqueue this rank's samples into the device sum buffer (no sync)
    finish()       -> reduce to rank 0, scale by 1/spp; rank 0 returns the HDR frame tensor
    """

    def __init__(self, dev_scene, cam: np.ndarray, rank: int = 0, world: int = 1, device_index: int = 0,
                 seed: int = 0xB200, variant: int = 0, frame=None, ldr=None):
        """`frame` / `ldr`: optional preallocated device tensors (H x W x 3 float32 / int32) to reuse
        across frames instead of allocating per renderer."""
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
        self.frame = frame if frame is not None else torch.empty((self.h, self.w, 3), dtype=torch.float32, device=self.device)
        self.ldr = ldr if ldr is not None else torch.empty((self.h, self.w, 3), dtype=torch.int32, device=self.device)

    def render_sum(self, want_stats: bool = False):
        from . import capi
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        return self.scene.render_device(self.cam, self.frame.data_ptr(), stream=stream, seed=self.seed,
                                        sample_offset=self.first, sample_count=self.count, variant=self.variant,
                                        flags=capi.FLAG_SUM, want_stats=want_stats)

    def finish(self, tonemap: bool = True):
        from . import capi
        reduce_frames(self.frame, dst=0)
        if self.rank != 0:
            return None
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        capi._check(capi.lib().b200rt_finalize_device(self.frame.data_ptr(), self.h * self.w, 1.0 / self.spp,
                                                      self.ldr.data_ptr() if tonemap else None, 0,
                                                      self.device.index, stream))
        return self.frame
