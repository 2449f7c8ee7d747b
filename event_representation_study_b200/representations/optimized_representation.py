"""Mirror of representations/optimized_representation.py (reference :14, :86-134): ERGO-12."""
from .representation_search.mixed_density_event_stack import MixedDensityEventStack

N_CHANNELS = 12


def get_optimized_representation(reshaped_return_data, num_events, height, width, _scale=None):
    window_indexes = [0, 3, 2, 6, 5, 6, 2, 5, 1, 0, 4, 1]
    functions = ["polarity", "timestamp_neg", "count_neg", "polarity", "count_pos", "count", "timestamp_pos", "count_neg",
                 "timestamp_neg", "timestamp_pos", "timestamp", "count"]
    aggregations = ["variance", "variance", "mean", "sum", "mean", "sum", "mean", "mean", "max", "max", "max", "mean"]
    stacking_type = ["SBN", "SBT"][0]
    transformation = MixedDensityEventStack(N_CHANNELS, num_events, height, width, (window_indexes, functions, aggregations), stacking_type)
    return transformation.stack(reshaped_return_data, _scale)
