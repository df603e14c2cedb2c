"""bayeslands_b200 - B200-native hot path of intelligentEarth/Bayeslands.

The pyBadlands forward model every MCMC proposal runs (`bl_mcmc.blackBox`) plus the likelihood
(`bl_mcmc.likelihoodFunc`), batched over proposals, as hand-written sm_100a CUDA kernels behind a
C-ABI (`include/bayeslands_b200.h`).  Host code is Python and keeps the reference's
`Model.load_xml/run_to_time` and `bayeslands_mcmc.blackBox/likelihoodFunc` surface.
"""
__version__ = "0.1.0"
