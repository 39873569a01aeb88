"""B200-native wavefront path tracer for CudaTracerLib's traversal + path-tracing hot path.

Public surface: `Scene`, `PathTracer`, `WavefrontPathTracer` (Tracer-shaped mirrors of the reference plugin API) over the
C ABI in include/ctl_b200.h (hand-written sm_100a CUDA in csrc/).  No CPU fallback.
"""
from .distributed import DistributedFrame, DistributedPasses, tiles_of_rank, TILE  # noqa: F401
from .api import (PathTracer, WavefrontPathTracer, Scene, SceneView, Material, ImagePipeline, VARIANCE_DTYPE, lib, generate_sample_tables, generate_sample_tables_n, traversal_bytes, tile_owner,  # noqa: F401
                  RAY_DTYPE, RESULT16_DTYPE, TRACE_RESULT_DTYPE, PIXEL_DTYPE, LIB_PATH)
