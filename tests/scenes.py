"""Scene definitions live with the host-side code (mesoengine_b200/scenes.py); re-exported for the tests."""
from mesoengine_b200.scenes import *  # noqa: F401,F403
from mesoengine_b200.scenes import REF_SPHERE, grid_center_world, orbit_eyes, sphere_scene, terrain_scene  # noqa: F401
