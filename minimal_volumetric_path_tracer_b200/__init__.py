"""minimal_volumetric_path_tracer_b200 -- B200-native drop-in for the per-pixel radiance loop of
gabo99cas/minimal_volumetric_path_tracer (src/rt.cpp:767-805).

The product is libvpt_b200.so (hand-written CUDA for sm_100a behind the C-ABI of include/vpt.h).  This package is the
thin Python host mirror of that boundary: `api` (ctypes bindings, Params / Scene helpers, render calls), `distributed`
(one process per GPU, sample / tile sharding, one NCCL reduce of the HDR buffers) and `cli` (the reference's
`rt <spp>` -> image.ppm surface).  There is no CPU compute path: importing works anywhere, rendering needs a CUDA device
and the built library (run `python -m minimal_volumetric_path_tracer_b200.build` or `__graft_entry__.build()`).
"""
from .api import (  # noqa: F401
    LIB_PATH, VptError, Params, Stats, Sphere, default_params, default_scene, load_scene, scene_from_rows, scene_to_rows, load_library,
    render, render_device, render_multi, render_into, PinnedFrame, unit, unit_strides, philox, measure_fp32_peak, tonemap, write_ppm, write_pfm, read_pfm, device_count, version,
    METHOD_FREE_FLIGHT, METHOD_EQUIANGULAR, METHOD_MIS, METHOD_RAYMARCH, METHOD_MIS_DISTANCE, METHOD_VOLUME_SPHERES, PRECISION_FP32, PRECISION_FP64_REF, OUTPUT_SUM, OUTPUT_MEAN,
    QUIRKS_NONE, QUIRKS_REFERENCE, QUIRK_R0_FALLTHROUGH, QUIRK_EXACT_VISIBILITY, UNIT, KERNEL_AUTO, KERNEL_MEGA, KERNEL_MEGA_SCAN, KERNEL_WAVEFRONT, KERNEL_WAVEFRONT_SM, KERNEL_WAVEFRONT_HBM,
)
