class AgentEnvWrapper:  # type-name only (planning/utils.py:7)
    def __init__(self, env, agent_fn):
        self.env = env
        self.agent_fn = agent_fn
