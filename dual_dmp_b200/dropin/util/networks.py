"""drop-in for reference util/networks.py -> dual_dmp_b200.util.networks"""
from dual_dmp_b200.util.networks import *  # noqa: F401,F403
from dual_dmp_b200.util.networks import GCNConv, NormalNet, PosNet  # noqa: F401
