from .pysynth import *  # noqa: F401,F403
