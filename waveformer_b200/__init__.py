"""waveformer_b200 - B200-native (sm_100a) implementation of WaveFormer's 3D-segmentation hot path.

``waveformer_b200.network_models`` mirrors the reference's ``network_models`` API; ``waveformer_b200.inferers`` is the
re-hosted sliding-window inferer; ``waveformer_b200.ops`` are the torch-level operators over the C ABI declared in
``include/waveformer_b200.h``.  Everything computes on CUDA; there is no CPU fallback.
"""
__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not import torch-heavy modules (python -m waveformer_b200.build)
    if name == "prepare_inference":
        from .precision import prepare_inference
        return prepare_inference
    raise AttributeError(name)
