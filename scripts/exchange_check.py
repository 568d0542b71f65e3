"""Under torchrun (>= 2 GPUs): the fused peer-memory exchange vs NCCL reduce + finalize.

Renders one small frame through FrameRenderer with both exchanges and compares the results on rank 0
(bit-equal at world size 2; FP32 summation order beyond), then times finish() alone for both at
1920x1080 and 3840x2160 with CUDA events (max over ranks).  Writes one JSON object to --out.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import cpp_raytracer_b200 as rt
    from cpp_raytracer_b200 import dist as rtdist, scene_io

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    scene = scene_io.load_scene(os.path.join(root, "tests", "golden", "rtow_lights.scene.gz"))
    res = {"world": world}

    w, h, spp = 250, 141, 16                      # ragged on purpose: 35250 pixels, not a multiple of 1024 or 4
    cam = rt.camera_with(scene.camera, image_w=w, image_h=h, spp=spp)
    with rt.DeviceSceneHandle(scene, device=local) as ds:
        a = rtdist.FrameRenderer(ds, cam, rank, world, local)
        a.render_sum()
        fa = a.finish(tonemap=True)
        peers = rtdist.PeerFrames(h, w, dev)
        b = rtdist.FrameRenderer(ds, cam, rank, world, local, peers=peers)
        b.render_sum()
        fb = b.finish(tonemap=True)
        torch.cuda.synchronize()
        if rank == 0:
            res["hdr_equal"] = bool(torch.equal(fa, fb)) if world == 2 else bool(torch.allclose(fa, fb, rtol=1e-5, atol=1e-6))
            res["hdr_max_abs_diff"] = float((fa - fb).abs().max())
            res["ldr_equal"] = bool(torch.equal(a.ldr, b.ldr)) if world == 2 else bool((a.ldr - b.ldr).abs().max() <= 1)
            one = rt.DeviceSceneHandle(scene, device=local)
            full, _ = one.render(cam)             # all samples on one GPU
            res["vs_single_gpu_max_rel"] = float(np.max(np.abs(fb.cpu().numpy() - full) / (np.abs(full) + 1e-6)))
            one.close()
        del peers, b

    def time_finish(fr):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
        for _ in range(3):
            fr.finish(tonemap=True)
        dist.barrier(); torch.cuda.synchronize()
        for s, e in ev:
            fr.frame.normal_(1.0, 0.1)            # keeps ranks loosely in step and the data fresh
            s.record(); fr.finish(tonemap=True); e.record()
        torch.cuda.synchronize()
        t = torch.tensor([float(np.median([s.elapsed_time(e) for s, e in ev]))], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    for (w, h) in ((1920, 1080), (3840, 2160)):
        cam = rt.camera_with(scene.camera, image_w=w, image_h=h, spp=64)
        with rt.DeviceSceneHandle(scene, device=local) as ds:
            nccl = rtdist.FrameRenderer(ds, cam, rank, world, local)
            nccl.frame.zero_()
            t_nccl = time_finish(nccl)
            peers = rtdist.PeerFrames(h, w, dev)
            peer = rtdist.FrameRenderer(ds, cam, rank, world, local, peers=peers)
            peer.frame.zero_()
            t_peer = time_finish(peer)
            res[f"finish_ms_{w}x{h}"] = {"nccl_reduce_plus_finalize": t_nccl, "peer_fused": t_peer,
                                          "frame_MB": w * h * 12 / 1e6}
            del peers, peer
    if rank == 0:
        print(json.dumps(res), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
