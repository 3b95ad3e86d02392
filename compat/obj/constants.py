from constants import *  # noqa: F401,F403  (`from obj.constants import ...` spelling of the reference)
