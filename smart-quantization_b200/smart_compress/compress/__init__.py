from .base import CompressionAlgorithmBase
from .bf16 import BF16
from .fp8 import FP8
from .fp16 import FP16
from .fp32 import FP32
from .s2fp8 import S2FP8
from .smart import SmartFP

# name -> class, as the reference's --compress flag maps them (smart_compress/util/train.py:119-126)
ALGORITHMS = dict(bf16=BF16, fp8=FP8, fp16=FP16, fp32=FP32, s2fp8=S2FP8, smart=SmartFP)

__all__ = ["CompressionAlgorithmBase", "BF16", "FP8", "FP16", "FP32", "S2FP8", "SmartFP", "ALGORITHMS"]
