from .quadruped_parkour import QuadrupedParkourEnv  # noqa: F401
from .humanoid_dancing import HumanoidDancingEnv  # noqa: F401
from .humanoid_soccer import HumanoidSoccerEnv  # noqa: F401
from .bipedal_rescue import BipedalRescueEnv  # noqa: F401
from .humanoid_construction import HumanoidConstructionEnv  # noqa: F401
from .humanoid_martial_arts import HumanoidMartialArtsEnv  # noqa: F401
from .robotic_arm_assembly import RoboticArmAssemblyEnv  # noqa: F401
