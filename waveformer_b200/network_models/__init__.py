"""Drop-in for the reference's ``network_models`` package (``network_models/__init__.py:10-61``): same names, same
constructor and forward signatures, same ``state_dict`` keys - computed by the B200 kernels in ``waveformer_b200``."""
from .network_backbone import Waveformer, create_waveformer, ProjectionHead, ChannelCalibration
from .waveformer import MultiscaleTransformer
from .wave_helper import Block, PatchMerging, PatchMergingV2, CCF_FFN, WaveletTransform3D, ProjectionUpsample
from .legacy import Mlp, DWConv, OverlapPatchEmbed, PatchEmbed, PosCNN
from .idwt_upsample import UnetrIDWTBlock as IDWTBlock, HFRefinementRes
from .attention import Attention

__version__ = "1.0.0"

__all__ = [
    "Waveformer", "create_waveformer", "ProjectionHead", "ChannelCalibration", "MultiscaleTransformer", "Block",
    "PatchMerging", "PatchMergingV2", "CCF_FFN", "Mlp", "WaveletTransform3D", "DWConv", "OverlapPatchEmbed",
    "PatchEmbed", "PosCNN", "ProjectionUpsample", "IDWTBlock", "HFRefinementRes", "Attention",
]
