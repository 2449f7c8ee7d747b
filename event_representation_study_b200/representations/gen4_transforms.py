"""Mirror of representations/gen4_transforms.py::get_item_transform (reference :12-83): the 1 Mpx variant has the same
branches as the Gen1 one, without the time_window argument."""
from .gen1_transforms import get_item_transform as _gen1


def get_item_transform(reshaped_return_data, representation_name, transform, height, width, num_events):
    return _gen1(reshaped_return_data, representation_name, transform, height, width, num_events, None)
