"""``render_flyaround`` under the reference's module path
(/root/reference/holo_diffusion/utils/render_utils/flyaround.py:45-297) for the sampling mode ``generate_samples.py``
drives: same keyword signature, same loop (dummy batch -> trajectory -> per-pose ``model(**batch, evaluation_mode=...,
voxel_features=...)`` -> images from preds -> videos), with the per-view post-processing on the GPU
(``holo_depth_image`` / ``holo_shade_depth`` / ``holo_frame_u8``) and ONE device-to-host copy of all 8-bit frames per
key at the end instead of a blocking ``.cpu()`` per key per view.

Reconstruction mode (``sample_mode=False``, flyaround.py:148-171): the sequence's frames come from `dataset`
(``sequence_indices_in_order`` + ``__getitem__``; collated with the frame type's own ``collate`` when it has one), the
source views are drawn with the reference's seeded ``randperm``, and the voxel grid comes from the view-pooling encoder
(``holo_diffusion_b200/encoder.py``).  The reference re-encodes the SAME source views for every pose of the trajectory
(only camera 0, the target, changes and it is not a source); here the grid is encoded once and passed to the remaining
poses as ``voxel_features`` -- same images, one encoder pass per fly-around instead of one per frame.
"""
from __future__ import annotations

import logging
import math
import os
from dataclasses import dataclass, fields
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

import holo_diffusion_b200 as _b200
from holo_diffusion_b200 import ops
from holo_diffusion_b200.cameras import PerspectiveCameras, get_simple_360_camera_trajectory  # noqa: F401

logger = logging.getLogger(__name__)

try:
    from pytorch3d.implicitron.dataset.dataset_base import FrameData
except ImportError:

    @dataclass
    class FrameData:   # the mapping protocol `model(**batch)` relies on, and .to(device)
        frame_number: Any = None
        sequence_category: Any = None
        image_rgb: Any = None
        camera: Any = None
        fg_probability: Any = None
        mask_crop: Any = None
        depth_map: Any = None
        sequence_name: Any = None
        frame_timestamp: Any = None

        def keys(self):
            return [f.name for f in fields(self)]

        def __getitem__(self, k):
            return getattr(self, k)

        def to(self, device):
            out = FrameData(**{k: self[k] for k in self.keys()})
            for k in out.keys():
                v = getattr(out, k)
                if hasattr(v, "to"):
                    setattr(out, k, v.clone().to(device) if hasattr(v, "clone") else v.to(device))
            return out


def _clone_cameras(c: PerspectiveCameras) -> PerspectiveCameras:
    return PerspectiveCameras(c.focal_length.clone(), c.principal_point.clone(), c.R.clone(), c.T.clone())


if not hasattr(PerspectiveCameras, "clone"):
    PerspectiveCameras.clone = _clone_cameras


def _get_dummy_test_batch_for_sampling(batch_size: int, device=torch.device("cpu")) -> FrameData:
    """flyaround.py:365-384: identity cameras that the loop overwrites pose by pose."""
    cam = PerspectiveCameras(R=torch.eye(3)[None].repeat(batch_size, 1, 1), T=torch.zeros(batch_size, 3),
                             principal_point=torch.ones(batch_size, 2), focal_length=torch.zeros(batch_size, 2))
    return FrameData(camera=cam.to(device))


def _collate_frames(frames: Sequence[Any]) -> Any:
    """FrameData.collate for a list of single frames (flyaround.py:386-396 goes through a DataLoader for it)."""
    collate = getattr(type(frames[0]), "collate", None)
    if collate is not None:
        return collate(list(frames))
    out = FrameData()
    for k in out.keys():
        vals = [getattr(f, k, None) for f in frames]
        if all(v is None for v in vals):
            continue
        if k == "camera":
            cat = lambda a: torch.cat([getattr(c, a) for c in vals], 0)   # noqa: E731  (one camera per frame, batch dim 1)
            out.camera = PerspectiveCameras(cat("focal_length"), cat("principal_point"), cat("R"), cat("T"))
        elif torch.is_tensor(vals[0]):
            setattr(out, k, torch.stack(vals, 0))
        else:
            setattr(out, k, list(vals))
    return out


def _stack_images(ims: torch.Tensor, size: Optional[Tuple[int, int]]) -> torch.Tensor:
    """The source views as one mosaic, row-major tiles on the smallest square grid that holds them, black where the grid
    is not filled (what flyaround.py:485-503 builds with nested concatenations)."""
    n, c, h, w = ims.shape
    side = int(math.ceil(math.sqrt(n)))
    canvas = ims.new_zeros(side * side, c, h, w)
    canvas[:n] = ims
    mosaic = canvas.view(side, side, c, h, w).permute(2, 0, 3, 1, 4).reshape(c, side * h, side * w)
    if size is not None:
        mosaic = torch.nn.functional.interpolate(mosaic[None], size=size, mode="bilinear")[0]
    return mosaic.clamp(0.0, 1.0)


def _images_from_preds(preds: Dict[str, Any], extract_keys: Sequence[str] = (
        "image_rgb", "images_render", "fg_probability", "masks_render", "depths_render", "depth_map",
        "_all_source_images")) -> Dict[str, torch.Tensor]:
    """flyaround.py:422-482 on the device: every value is a (1, 3, H, W) float tensor that stays on the GPU."""
    imout: Dict[str, torch.Tensor] = {}
    for k in extract_keys:
        if k == "_all_source_images":
            if preds.get("image_rgb") is None:
                continue   # sampling mode has no source images
            imout[k] = _stack_images(preds["image_rgb"][1:].detach(), None)[None]
            continue
        if k == "_shaded_depth_render" and ("normals_render" in preds or "depths_render" in preds):
            d, m = preds["depths_render"][0, 0].contiguous(), preds["masks_render"][0, 0].contiguous()
            cam = preds["camera"]
            H, W = d.shape
            kk = int(math.ceil(0.005 * math.sqrt(H ** 2 + W ** 2)))   # smoothing_kernel_size=0.005
            if "_camera_intrinsics_host" in preds:   # (fx, fy), (px, py) kept on the host by render_flyaround: no sync
                f, p = preds["_camera_intrinsics_host"]
            else:
                f, p = cam.focal_length[0].tolist(), cam.principal_point[0].tolist()
            v, _ = ops.shade_depth(d, m, f, p, kk, mask_thr=0.5, depth_thr=1e-2, material="medium", bg=(1.0, 1.0, 1.0))
            v = v[None]
        else:
            if k not in preds or preds[k] is None:
                logger.debug(f"cant show {k}")
                continue
            v = preds[k]
            if k.startswith("depth"):
                m = preds["masks_render"]
                if m.shape[2:] != v.shape[2:]:
                    raise NotImplementedError("mask and depth of different sizes")
                v = ops.depth_image(v[0, 0].contiguous(), m[0, 0].contiguous())[0][None]
        if v.shape[1] == 1:
            v = v.expand(-1, 3, -1, -1)
        imout[k] = v
    return imout


def _generate_prediction_videos(preds: List[Dict[str, torch.Tensor]], sequence_name: str, predicted_keys: Sequence[str],
                                fps: int = 20, video_path: str = "/tmp/video", video_frames_dir: Optional[str] = None,
                                resize: Optional[Tuple[int, int]] = None) -> Dict[str, str]:
    """flyaround.py:558-610: 8-bit frames (clip, resize) per key.  The frames of a key are packed on the GPU into one
    (n, h, w, 3) uint8 tensor and copied to the host once; they are written as ``<video_path>_<name>_<key>.npy`` and,
    when an ``ffmpeg`` binary is on PATH, encoded to ``.mp4`` like pytorch3d's VideoWriter does."""
    import shutil
    import subprocess
    os.makedirs(os.path.dirname(video_path) or ".", exist_ok=True)
    written = {}
    for k in predicted_keys:
        if k not in preds[0]:
            continue
        H, W = preds[0][k].shape[-2:]
        h, w = (H, W) if resize is None else (int(resize[0]), int(resize[1]))
        dev = preds[0][k].device
        frames = torch.empty(len(preds), h, w, 3, dtype=torch.uint8, device=dev)
        for i, p in enumerate(preds):
            ops.frame_u8(p[k][0].contiguous(), (h, w), frames[i])
        if frames.is_cuda:
            host = torch.empty(frames.shape, dtype=torch.uint8, pin_memory=True)
            host.copy_(frames, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:   # host-logic tests
            host = frames
        out = f"{video_path}_{sequence_name}_{k}"
        np.save(out + ".npy", host.numpy())
        written[k] = out + ".npy"
        if shutil.which("ffmpeg"):
            cmd = ["ffmpeg", "-y", "-loglevel", "error", "-f", "rawvideo", "-pix_fmt", "rgb24", "-s", f"{w}x{h}", "-r",
                   str(fps), "-i", "-", "-pix_fmt", "yuv420p", out + ".mp4"]
            if subprocess.run(cmd, input=host.numpy().tobytes()).returncode == 0:
                written[k] = out + ".mp4"
        logger.info(f"Generated {written[k]}.")
    return written


def render_flyaround(dataset, sequence_name: str, model: torch.nn.Module, output_video_path: str,
                     output_video_name: Optional[str] = None, n_flyaround_poses: int = 40, fps: int = 20,
                     trajectory_type: str = "circular_lsq_fit", max_angle: float = 2 * math.pi,
                     trajectory_scale: float = 1.1, scene_center: Tuple[float, float, float] = (0.0, 0.0, 0.0),
                     up: Tuple[float, float, float] = (0.0, -1.0, 0.0),
                     camera_elevation: float = -30.0 * (2 * math.pi / 360), camera_focal_length: float = 3.2,
                     hemispherical_radius: float = 10, traj_offset: float = 0.0, n_source_views: int = 9,
                     visdom_show_preds: bool = False, visdom_environment: str = "render_flyaround",
                     visdom_server: str = "http://127.0.0.1", visdom_port: int = 8097, num_workers: int = 10,
                     device: Union[str, torch.device] = "cuda", seed: Optional[int] = None,
                     video_resize: Optional[Tuple[int, int]] = None, output_video_frames_dir: Optional[str] = None,
                     sample_mode: bool = False, progressive_sampling_steps_per_render: int = -1,
                     visualize_preds_keys: Sequence[str] = ("images_render", "masks_render", "depths_render",
                                                            "_all_source_images"),
                     save_voxel_features: bool = False):
    if visdom_show_preds:
        raise NotImplementedError("visdom output is not part of the built path")
    if seed is None:
        seed = hash(sequence_name)
    if sample_mode:
        batch = _get_dummy_test_batch_for_sampling(n_source_views + 1, device="cpu")
    else:   # reconstruction mode (flyaround.py:152-171)
        logger.info(f"Loading all data of sequence '{sequence_name}'.")
        seq_idx = list(dataset.sequence_indices_in_order(sequence_name))
        with torch.random.fork_rng():   # sample the source views reproducibly
            torch.manual_seed(seed)
            source_views_i = torch.randperm(len(seq_idx))[:n_source_views]
        # the first, dummy view gets replaced with the target camera
        source_views_i = torch.nn.functional.pad(source_views_i, [1, 0])
        batch = _collate_frames([dataset[seq_idx[i]] for i in source_views_i.tolist()])
        assert all(batch.sequence_name[0] == sn for sn in batch.sequence_name)
    if trajectory_type.lower() != "simple_360":
        raise NotImplementedError("trajectories fitted to the training cameras (pytorch3d generate_eval_video_cameras) "
                                  "are not built: use 'simple_360' (generate_samples.py:46)")
    test_cameras = get_simple_360_camera_trajectory(max_angle, n_flyaround_poses, camera_elevation,
                                                    hemispherical_radius, up, camera_focal_length)
    EvaluationMode = _b200.EvaluationMode
    voxel_features = None
    if sample_mode and progressive_sampling_steps_per_render <= 0:
        voxel_features = model.sample_random_voxel_features()
    gen = model.sample_random_voxel_features_progressive() if progressive_sampling_steps_per_render > 0 else None
    encoded = None   # reconstruction: the grid of the (pose-independent) source views, encoded once
    preds_total = []
    for n in range(n_flyaround_poses):
        for k in ("R", "T", "focal_length", "principal_point"):   # the first batch camera becomes the target camera
            getattr(batch.camera, k)[0] = getattr(test_cameras[n], k)[0]
        net_input = batch.to(device)
        with torch.no_grad():
            if gen is not None:
                for _ in range(progressive_sampling_steps_per_render):
                    try:
                        voxel_features = next(gen)
                    except StopIteration:
                        break
            kw = {k: net_input[k] for k in net_input.keys()}
            kw.update(evaluation_mode=EvaluationMode.EVALUATION, voxel_features=voxel_features)
            if voxel_features is None and kw.get("image_rgb") is not None:
                impl = _b200.renderer.impl_of(model)
                names = kw.get("sequence_name")
                pose_free = names is not None and sum(1 for s_ in names if s_ == names[0]) > 1   # view 0 is not a source
                if encoded is None or not pose_free:
                    encoded = impl.encode_views(image_rgb=kw["image_rgb"], camera=kw["camera"],
                                                fg_probability=kw.get("fg_probability"), mask_crop=kw.get("mask_crop"),
                                                sequence_name=names)
                kw.update(image_rgb=None, voxel_features=encoded)
            preds = model(**kw)
            assert all(k not in preds for k in net_input.keys())
            preds.update({k: net_input[k] for k in net_input.keys()})
            preds["_camera_intrinsics_host"] = (test_cameras[n].focal_length[0].tolist(),
                                                test_cameras[n].principal_point[0].tolist())
            preds_total.append(_images_from_preds(preds, extract_keys=visualize_preds_keys))
    if output_video_name is None:
        output_video_name = sequence_name
    logger.info(f"Exporting videos for sequence {sequence_name} ...")
    written = _generate_prediction_videos(preds_total, sequence_name=output_video_name, fps=fps,
                                          video_path=output_video_path, resize=video_resize,
                                          video_frames_dir=output_video_frames_dir, predicted_keys=visualize_preds_keys)
    if voxel_features is not None and save_voxel_features:
        logger.info(f"Saving voxel features for sequence {sequence_name} ...")
        output_directory = "/".join(output_video_path.split("/")[:-1])
        torch.save(voxel_features, os.path.join(output_directory, f"{sequence_name}_voxel_features.pth"))
    return written
