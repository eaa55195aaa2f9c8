"""thrifty_b200 -- B200-native implementation of Thrifty's `detect` hot path.

Host-side mirror of the reference's interface for this path (`detect`, `block_data`,
`toads_data`, `settings`, `setting_parsers`) over hand-written sm_100a CUDA kernels reached
through a C ABI (include/thrifty_b200.h).  Importing the package does not load the shared
object; constructing a detector does, and raises if it is missing (no CPU fallback).
"""

__version__ = "0.1.0"
