from oracle.vec_env import FlattenObservations, RecordEpisodeStatistics, StackEnv  # noqa: F401


class RecordVideo:  # name only
    def __init__(self, env, *a, **k):
        self.env = env
