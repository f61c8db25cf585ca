"""berserkerquake2_b200 -- B200-native alias-model lerp + stencil-shadow-volume geometry.

Host-side Python mirror of the C ABI in ``include/bq2_gpu.h`` (which itself replaces the reference's
``R_CalcAliasFrameLerp`` / ``R_CalcAliasFrameLerpShadow`` / ``R_CalcAliasFrameShadowVolume`` /
``R_DrawAliasShadowVolume`` / ``R_DrawMD3AliasShadowVolume``, main.c:63510-64038).  Everything that
computes is in ``libbq2gpu.so`` (hand-written sm_100a CUDA + C); this package only marshals numpy
buffers through ctypes.  There is no CPU fallback: without the built library importing
:mod:`berserkerquake2_b200._capi` raises, and without a B200 ``Context()`` raises ``Bq2Error``.
"""
from ._capi import (Bq2Error, ENTITY_DTYPE, PAIR_DTYPE, WORLD_ENTITY_DTYPE, WORLD_LIGHT_DTYPE,  # noqa: F401
                    WANT_SHADOW, WANT_NTB, WANT_TSLIGHT, WANT_NOLERPOUT, WANT_TABLES, MODEL_SMOOTHVECS, lib, lib_path)
from .context import Context, FrameResult  # noqa: F401
from . import synth, host, shard  # noqa: F401

__all__ = ["Context", "FrameResult", "Bq2Error", "synth", "host", "shard", "ENTITY_DTYPE", "PAIR_DTYPE",
           "WORLD_ENTITY_DTYPE", "WORLD_LIGHT_DTYPE", "WANT_SHADOW", "WANT_NTB", "WANT_TSLIGHT",
           "WANT_NOLERPOUT", "WANT_TABLES", "MODEL_SMOOTHVECS", "lib", "lib_path"]
