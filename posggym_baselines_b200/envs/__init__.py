from posggym_baselines_b200.envs.driving import BatchedDriving, DrivingModel  # noqa: F401
from posggym_baselines_b200.envs.env import Env, make  # noqa: F401
from posggym_baselines_b200.envs.lbf import BatchedLBF, LBFModel, LevelBasedForagingModel  # noqa: F401
from posggym_baselines_b200.envs.predator_prey import BatchedPredatorPrey, PredatorPreyModel  # noqa: F401
