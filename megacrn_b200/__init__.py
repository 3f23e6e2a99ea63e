"""megacrn_b200 -- B200-native (sm_100a) implementation of the MegaCRN hot path.

Public surface mirrors the reference's ``model/MegaCRN.py``: ``MegaCRN`` (nn.Module),
``print_params``; plus ``ddp`` helpers for batch-sharded training.
"""
from .MegaCRN import ADCRNN_Decoder, ADCRNN_Encoder, AGCN, AGCRNCell, MegaCRN, print_params  # noqa: F401

__all__ = ["MegaCRN", "AGCN", "AGCRNCell", "ADCRNN_Encoder", "ADCRNN_Decoder", "print_params"]
