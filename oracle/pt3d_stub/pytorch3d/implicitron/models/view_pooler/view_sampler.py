"""Stand-in leaf (TEST INFRASTRUCTURE ONLY), restated from memory of pytorch3d 0.7.4 -- unpinned."""
import torch


def cameras_points_cartesian_product(camera, pts):
    """[camera[0] x every point batch, camera[1] x ..., ...] and the points repeated alongside."""
    n_cameras, batch_pts = camera.R.shape[0], pts.shape[0]
    pts_rep = pts.repeat(n_cameras, *([1] * (pts.ndim - 1)))
    idx = torch.arange(n_cameras)[:, None].expand(n_cameras, batch_pts).reshape(batch_pts * n_cameras)
    return camera[idx.tolist()], pts_rep
