"""pygradflow_b200 -- B200-native batched implementation of pygradflow's inner Newton/KKT path.

Importing the package is cheap (no torch, no CUDA); the CUDA library is loaded on first use by
``pygradflow_b200.native`` and raises if it is missing -- there is no CPU fallback.
"""

__version__ = "0.1.0"
