"""b200-gamma: ctypes mirror of the B200-native gamma hot path (libgamma.so).

    index.GammaIndex   IndexModel seam (include/gamma_b200_index.h): FLAT / IVFFLAT / IVFPQ on one GPU
    engine.GammaEngine gamma's own C-ABI (include/gamma_api.h), driven like the Go partition server does
    wire               flatbuffers / protobuf codecs of the payloads that cross that ABI
    synth              synthetic SIFT- and embedding-shaped data for the parity tests and bench.py

There is no CPU implementation behind these wrappers: importing works anywhere, creating an index or an
engine needs a CUDA device and the in-tree extension built by ``__graft_entry__.build()``.
"""
