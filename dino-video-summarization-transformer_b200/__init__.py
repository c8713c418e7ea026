"""dvst_b200 - B200-native (sm_100a) implementation of the Dino-Video-Summarization-Transformer scoring path.

Host side mirrors the reference's Python interfaces; all device math lives in csrc/libdvst.so (C ABI, include/dvst.h).
"""
__version__ = "0.1.0"
