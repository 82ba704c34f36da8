"""Synthetic test meshes: the generators live in fcl_b200/workloads.py (bench.py uses them too)."""
from fcl_b200.workloads import *  # noqa: F401,F403
from fcl_b200.workloads import box_mesh, heightfield, noisy_sphere, random_soup, serial_chain_poses, uv_sphere  # noqa: F401
