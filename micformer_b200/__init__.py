"""micformer_b200 -- B200-native (sm_100a) implementation of the MicFormer dual-stream forward/backward hot path.

Public surface (mirrors the reference's modules):
    micformer_b200.models.MICFormer_self.Head / MicFormer / CrossTransformerBlock3D / ...
    micformer_b200.models.STN.SpatialTransformer
    micformer_b200.loss.dice.MDiceLoss
"""
__version__ = "0.1.0"
