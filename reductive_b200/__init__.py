"""reductive_b200 — B200-native (sm_100a) product-quantization hot path behind reductive's quantizer API.

Importing this package loads reductive_b200/lib/libreductive_b200.so (hand-written CUDA kernels + C ABI,
include/reductive_b200.h).  There is no CPU fallback: a missing library is an ImportError and a missing GPU
is a NoDeviceError at the first compute call.
"""
from ._cabi import (  # noqa: F401
    ENCODE_AUTO, ENCODE_EXACT, ENCODE_TENSOR, PROJECT_AUTO, PROJECT_EXACT, PROJECT_TENSOR, ConstructRng, CudaError, IncorrectNAttempts, IncorrectNIterations,
    IncorrectNSubquantizerBits, IncorrectNumberSubquantizers, NoDeviceError, NSubquantizersOutsideRange,
    ReductiveError, ReductivePanic, kernel_launch_count, set_encode_algo, set_kmeans_update,
    set_project_algo,
)
from .kmeans import (  # noqa: F401
    KMeans, NIterationsCondition, RandomInstanceCentroids, kmeans_iteration, kmeans_with_centroids,
)
from .opq import GaussianOpq, Opq, bucket_eigenvalues  # noqa: F401
from .pq import Pq, QuantizeVector, Reconstruct, TrainPq, check_quantizer_invariants  # noqa: F401
from .storage import QuantizedArray  # noqa: F401

__all__ = [
    "Pq", "Opq", "GaussianOpq", "TrainPq", "QuantizeVector", "Reconstruct", "KMeans", "NIterationsCondition",
    "RandomInstanceCentroids", "kmeans_iteration", "kmeans_with_centroids", "ReductiveError", "ReductivePanic",
    "QuantizedArray",
]
