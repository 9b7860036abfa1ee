"""drop-in for reference util/mesh.py -> dual_dmp_b200.util.mesh"""
from dual_dmp_b200.util.mesh import *  # noqa: F401,F403
from dual_dmp_b200.util.mesh import Mesh  # noqa: F401
