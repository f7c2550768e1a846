"""nimblesm_b200 — B200-native hex8 explicit-dynamics path behind NimbleSM's Block / ModelData API.

csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/nsm_b200.h) -> lib/libnsm_b200.so
capi.py    ctypes mirror of the C ABI (no fallback: missing library or GPU raises)
deck.py    input-deck surface (nimble::Parser keys)
mesh.py    synthetic structured hex8 cubes, element partitions, shared-node tables
model.py   DataManager / ModelData / ExplicitTimeIntegrator sequencing on top of the C ABI
host/      C++ twin of the above for drop-in use from NimbleSM-style C++ code
"""
__version__ = "0.1.0"
