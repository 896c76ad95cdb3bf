"""team_b200 - B200-native drop-in for the TEAM/PROOF multimodal head.

Host side (Python, mirrors the reference's class surface) over a C-ABI shared library of
hand-written sm_100a CUDA kernels (``csrc/`` -> ``libteam_b200.so``).  There is no CPU or
PyTorch fallback: every op raises if the library or a B200 is missing.
"""
__version__ = "0.1.0"
