"""Device-side pieces of the PPO learner's batch post-processing."""
from posggym_baselines_b200.ppo.gae import compute_gae  # noqa: F401
