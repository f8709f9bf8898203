"""deflatedmlmc_schwinger_b200 -- B200-native (sm_100a CUDA behind a C ABI) hot path of the
deflated multilevel-Monte-Carlo Hutchinson estimator for tr(D^-1) of the 2-D Schwinger
Wilson-Dirac operator, behind the Python entry points of the reference
(main / gateway / examples / matrix / multigrid / stoch_trace / utils)."""
__version__ = "0.1.0"
