from lightning import *  # noqa: F401,F403  (`from obj.lightning import ...` spelling of the reference)
from lightning import Lightning  # noqa: F401
