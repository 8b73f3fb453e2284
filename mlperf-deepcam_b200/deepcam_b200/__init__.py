"""deepcam_b200 — B200-native kernels and engine behind the DeepCAM reference API.

Sub-modules: build (nvcc in-tree build), _lib (ctypes C-ABI binding), ops (tensor-level kernel wrappers),
convdesc (tap tables), backend (layer-level CUDA operations), engine (network forward/backward plan).
"""
__all__ = ["build", "_lib", "ops", "convdesc", "backend"]
