"""evrep-b200: B200-native event-representation engine (hand-written sm_100a CUDA behind a C ABI).

`batched`      the engine API: CSR-packed windows in, CUDA tensors out (one call per representation)
`representations`, `evlicious`, `tonic_compat`   drop-in mirrors of the reference's per-window interfaces
`synth`        seeded synthetic event streams
Importing the package loads event_representation_study_b200/lib/libevrep.so and fails loudly if it is absent.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is not built)

__version__ = "0.1.0"
