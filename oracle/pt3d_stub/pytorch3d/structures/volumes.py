class VolumeLocator:
    """Holds what HoloVoxelGridImplicitFunction.forward passes (holo_voxel_grid_implicit_function.py:204-209)."""

    def __init__(self, batch_size, grid_sizes, device, voxel_size, volume_translation=(0.0, 0.0, 0.0), align_corners=True):
        assert batch_size == 1 and align_corners
        self.grid_sizes, self.voxel_size, self.device = tuple(grid_sizes), voxel_size, device
