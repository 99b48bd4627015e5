from .pvder_env import PVDER  # noqa: F401
from .vec_env import PVDERVecEnv  # noqa: F401
