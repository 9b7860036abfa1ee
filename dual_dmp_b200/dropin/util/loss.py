"""drop-in for reference util/loss.py -> dual_dmp_b200.util.loss"""
from dual_dmp_b200.util.loss import *  # noqa: F401,F403
