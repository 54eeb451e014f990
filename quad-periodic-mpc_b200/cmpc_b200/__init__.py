"""Host-side Python bindings (ctypes over the C-ABI in include/cmpc_b200.h) and
synthetic workload generators for the B200 batched convex-MPC engine."""
