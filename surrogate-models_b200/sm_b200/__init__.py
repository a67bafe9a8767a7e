"""sm_b200 -- loader and thin Python plumbing over libsurrogate_b200.so (the sm_100a C-ABI library).

The reference-facing API lives in the sibling packages ``mylibs`` / ``myutils`` (same import paths as
juri117/surrogate-models); this package only binds the C-ABI (``_cabi``), owns device memory through
torch (``device``) and shards candidates / queries across ranks (``dist``).
"""
from . import _cabi  # noqa: F401
