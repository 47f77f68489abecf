from .base_model import *  # noqa: F401,F403
