from cube_map import *  # noqa: F401,F403  (`from obj.cube_map import ...` spelling of the reference)
from cube_map import CubeMap  # noqa: F401
