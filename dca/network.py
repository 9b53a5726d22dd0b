from dca_b200.network import *  # noqa: F401,F403
from dca_b200.network import AE_types  # noqa: F401
