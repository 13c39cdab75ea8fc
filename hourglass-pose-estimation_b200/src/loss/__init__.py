from .mse import *  # noqa: F401,F403
