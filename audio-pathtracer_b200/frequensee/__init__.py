"""frequensee -- Python host side of the B200-native FrequenSee propagation core.

`capi` binds the C-ABI (include/frequensee.h, CUDA only, no CPU fallback); `scenes` holds the
seeded procedural scenes of the benchmark configs; `component` mirrors the reference's
UFrequenSeeAudioComponent / UAudioRayTracingSubsystem / FFrequenSeeAudioReverbPlugin interface
for this path on top of the C-ABI.
"""
from . import capi, scenes  # noqa: F401
from .capi import Context, FrequenSeeError, default_config, load_float_array, save_float_array  # noqa: F401
