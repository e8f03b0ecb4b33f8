"""B200-native seriation MCMC sweep: CUDA kernels + C ABI (csrc/, libseriation_b200.so) and the
host-side mirror of the reference's driver interface (api.py).  Import through the top-level
``seriation_b200`` module (the directory name carries hyphens)."""
from .api import *  # noqa: F401,F403
from . import api  # noqa: F401
