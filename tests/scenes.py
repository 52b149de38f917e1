"""Scene definitions live with the host-side code (mesoengine_b200/scenes.py).  Loaded by file path, not through the
package: importing mesoengine_b200 loads libmeso_b200.so, which oracle-only users (CPU tests, bench.py --impl reference)
must not do."""
import importlib.util
import os

_spec = importlib.util.spec_from_file_location(
    "_meso_scenes", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mesoengine_b200", "scenes.py"))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
REF_SPHERE, grid_center_world, orbit_eyes, sphere_scene, terrain_scene = (
    _m.REF_SPHERE, _m.grid_center_world, _m.orbit_eyes, _m.sphere_scene, _m.terrain_scene)
