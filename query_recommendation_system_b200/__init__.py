"""B200-native LSH → top-k collaborative-filtering hot path of
wamuumu/query-recommendation-system, behind the reference's own Python surface
(`LSH`, `Recommender`).  Kernels: hand-written CUDA for sm_100a in
`csrc/`, exported through the C ABI of `include/qrs_b200.h` (libqrs_b200.so).
There is no CPU fallback."""
from ._lib import QrsError, LIB_PATH, load as load_library  # noqa: F401
from .lsh import LSH  # noqa: F401

__all__ = ["LSH", "QrsError", "load_library", "LIB_PATH"]
