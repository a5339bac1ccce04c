"""B200-native (sm_100a) implementation of smarc_navigation's one data-parallel
hot path: the ``auv_particle_filter`` measurement cycle (propagate -> raycast ->
beam likelihood -> normalise + resample), behind the reference filter's own
``predict`` / ``update`` / ``resample`` API.  See DESIGN.md."""
from .filter import ParticleFilter, parse_covariance   # noqa: F401
from ._lib import AuvpfError, DegenerateWeights        # noqa: F401

__all__ = ["ParticleFilter", "parse_covariance", "AuvpfError", "DegenerateWeights"]
