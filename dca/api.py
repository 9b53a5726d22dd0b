from dca_b200.api import *  # noqa: F401,F403
from dca_b200.api import dca  # noqa: F401
