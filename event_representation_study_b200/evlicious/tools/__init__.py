from .utils import events_to_voxel_grid, events_to_voxel_grid_cuda  # noqa: F401
