from posggym_baselines_b200.vector.vec_env import BatchedVecEnv, load_vec_env  # noqa: F401
