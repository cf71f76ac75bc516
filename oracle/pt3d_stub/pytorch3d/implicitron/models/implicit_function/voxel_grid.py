"""Stand-in leaves (TEST INFRASTRUCTURE ONLY): FullResolutionVoxelGrid.evaluate_world = VolumeLocator.world_to_local_coords
+ interpolate_volume, delegated to the oracle's restatement -- unpinned arithmetic."""
from oracle import render_oracle as ro


class VoxelGridBase:
    pass


class VoxelGridValuesBase:
    pass


class FullResolutionVoxelGridValues(VoxelGridValuesBase):
    def __init__(self, voxel_grid):
        self.voxel_grid = voxel_grid


class FullResolutionVoxelGrid(VoxelGridBase):
    def __init__(self, n_features: int = 1, **kw):
        self.n_features = n_features

    def evaluate_world(self, points, grid_values, locator):
        resol = locator.grid_sizes[0]
        local = ro.world_to_local(points.reshape(-1, 3), resol, locator.voxel_size * resol)
        return ro.sample_grid(grid_values.voxel_grid, local).reshape(*points.shape[:-1], -1)
