from oracle.vec_env import SyncVectorEnv  # noqa: F401
