"""tensorflow_ocr_b200 — B200-native (sm_100a) PixelLink / EAST per-pixel text-detection head.

Drop-in for the head of BowieHsu/tensorflow_ocr: same function names and signatures
as ``nets.model``, ``nets.model_vgg_16``, ``nets.pixellink``, ``tool.pixellink_fn`` and
``datasets.icdar``, backed by hand-written CUDA kernels behind the
C ABI in ``include/plhead.h`` (``libplhead.so``).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .head import DecodeConfig, LossConfig  # noqa: F401

__version__ = "0.1.0"
