from .utils import events_to_voxel_grid, events_to_voxel_grid_cuda  # noqa: F401
from . import filters  # noqa: F401,E402  (reference tools/__init__.py:3)
