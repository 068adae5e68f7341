"""sceneverse_b200 — B200 (sm_100a) implementation of SceneVerse's GPS hot path.

Host side mirrors the reference's operator/module interface; compute is hand-written CUDA behind
the C-ABI declared in include/*.h.  See DESIGN.md and INTEGRATION.md.
"""
__version__ = "0.1.0"
