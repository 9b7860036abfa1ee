"""drop-in for reference util/models.py -> dual_dmp_b200.util.models"""
from dual_dmp_b200.util.models import *  # noqa: F401,F403
