"""ivit_b200 -- B200-native (sm_100a) integer-only ViT inference operators.

A from-scratch implementation of the hot path of zkkli/I-ViT (the frozen INT8 forward through
``models/quantization_utils``): hand-written CUDA kernels behind a C ABI
(``include/ivit_b200.h`` / ``csrc/``), a Python mirror of the reference's operator classes
(``quantization_utils``), model graphs built on them (``deit``, ``swin``), a frozen parameter
pack (``pack``) and a fused whole-model executor (``engine``).  No CPU fallback anywhere.
"""
__version__ = "0.1.0"
