"""keynet_b200 -- B200-native (sm_100a) implementation of the keyed-layer forward path of visym/keynet.

Host-side Python mirrors the reference's public surface for that path (Keynet / PermutationKeynet /
KeyedLayer / KeyedSensor / SparseMatrix ...); all matrix construction and the forward SpMM run in
hand-written CUDA behind the C ABI of include/keynet_b200.h.  There is no CPU fallback.
"""
from .version import __version__  # noqa: F401
