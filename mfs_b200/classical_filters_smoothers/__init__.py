"""Mirror of the one member of ``mfs/classical_filters_smoothers`` that is on the hot path: the brute-force grid
filter (``brute_force.py``).  The Gaussian filters/smoothers and SMC baselines of that package are out of scope."""
from .brute_force import brute_force_filter

__all__ = ['brute_force_filter']
