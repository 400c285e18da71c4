"""waveformer_b200 - B200-native (sm_100a) implementation of WaveFormer's 3D-segmentation hot path.

``waveformer_b200.network_models`` mirrors the reference's ``network_models`` API; ``waveformer_b200.inferers`` is the
re-hosted sliding-window inferer; ``waveformer_b200.ops`` are the torch-level operators over the C ABI declared in
``include/waveformer_b200.h``.  Everything computes on CUDA; there is no CPU fallback.
"""
__version__ = "0.1.0"

from .precision import prepare_inference  # noqa: E402,F401
