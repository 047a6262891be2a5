"""mp3-enc-bsd_b200 — B200-native (sm_100a) front end + rate loop of the lieff/mp3-enc-bsd Layer III
encoder, behind the C ABI of libmp3gpu.so (include/mp3gpu.h).

The directory name contains '-' and '.', so it is imported through `mp3gpu_pkg.load()` at the repo
root, which registers it as the module `mp3enc_b200`.
"""
from . import host, segment, synth  # noqa: F401
from .host import Encoder, Mp3GpuError, load_library  # noqa: F401
