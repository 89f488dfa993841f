"""sse_b200: host-side mirror of StableSpectralElements.jl's Solver surface for the
B200-native semi-discrete residual (libsse_b200.so)."""
