from dca_b200.train import *  # noqa: F401,F403
from dca_b200.train import train, train_with_args  # noqa: F401
