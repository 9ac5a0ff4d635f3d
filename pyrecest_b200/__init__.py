"""pyrecest_b200 -- the particle-filter hot path of pyRecEst (predict / update / resample /
estimate on R^d, T^d, S^d) as hand-written sm_100a CUDA kernels behind the reference's Python API.

    from pyrecest_b200.filters import EuclideanParticleFilter, HypertoroidalParticleFilter
    from pyrecest_b200.distributions import GaussianDistribution, LinearDiracDistribution

The CUDA library (pyrecest_b200/lib/libpfb.so, C ABI in include/pfb.h) is mandatory; nothing
here falls back to the CPU."""
__version__ = "0.1.0"
