"""drop-in for reference util/datamaker.py -> dual_dmp_b200.util.datamaker"""
from dual_dmp_b200.util.datamaker import *  # noqa: F401,F403
from dual_dmp_b200.util.datamaker import Dataset, create_dataset, dataset_from_meshes  # noqa: F401
