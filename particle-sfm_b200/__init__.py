"""B200-native ParticleSfM optimisation hot paths (HP1 path-consistency trajectory
optimiser, HP2 global bundle adjustment).  See DESIGN.md.

The CUDA library is loaded lazily (first use); there is no CPU fallback: without the
built extension or without a CUDA device every compute call raises."""
__version__ = "0.1.0"
