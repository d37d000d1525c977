"""faqcs_b200 -- B200-native (sm_100a) implementation of the FaQCs v2.10 per-read
trim / filter / statistics hot path behind a C ABI (include/faqcs_b200.h).

``faqcs_b200.api``   ctypes mirror of the C ABI (Engine, Options, Stats)
``faqcs_b200.synth`` seeded synthetic FASTQ for the BASELINE workloads
``faqcs_b200.csrc``  CUDA kernels + the C-ABI implementation (built into libfaqcs_b200.so)
``faqcs_b200.host``  C++ command-line driver mirroring the FaQCs CLI
"""
from .api import Engine, FaqcsError, Options, Stats  # noqa: F401

__version__ = "0.1.0"
