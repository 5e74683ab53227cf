from .linear import *  # noqa: F401,F403
