"""Synthetic MPAS input files for the tests: the generator lives beside the synthetic meshes (mpassit_b200/mpas_files.py)."""
from mpassit_b200.mpas_files import *  # noqa: F401,F403
from mpassit_b200.mpas_files import START_TIME, VALID_TIME  # noqa: F401
