"""Mirror of the ev-licious operator API the reference's hot path uses: `Events` and `tools.events_to_voxel_grid[_cuda]`
(ev-licious/src/evlicious/io/utils/events.py:11-45, tools/utils.py:7-85)."""
from .io.utils.events import Events, TYPES  # noqa: F401
from . import tools  # noqa: F401
