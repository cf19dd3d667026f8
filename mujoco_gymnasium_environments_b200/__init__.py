"""B200-native batched physics-and-task engine for the Mujoco_Gymnasium_Environments tasks."""
__version__ = "0.1.0"
