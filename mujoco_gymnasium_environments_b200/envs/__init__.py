from .quadruped_parkour import QuadrupedParkourEnv  # noqa: F401
