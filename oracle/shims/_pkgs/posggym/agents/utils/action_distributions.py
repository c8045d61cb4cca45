class DiscreteActionDistribution:  # name only (search_policy.py:9)
    def __init__(self, probs, rng=None):
        self.probs = probs
