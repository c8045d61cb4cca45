from posggym_baselines_b200.envs.driving import BatchedDriving, DrivingModel  # noqa: F401
