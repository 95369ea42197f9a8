"""B200-native batched window decoders with the SlidingWindowDecoder Python API."""
from .decoders import bpgdg_decoder, bpgd_decoder, osd_window, BpOsdDecoder  # noqa: F401

__version__ = "0.1.0"
