"""posggym.agents stand-in: only Random-v0 exists (the zoo is out of scope)."""
from posggym.agents.policy import Policy, PolicyState, RandomPolicy  # noqa: F401


def make(policy_id: str, model, agent_id: str, **kwargs):
    if policy_id.split("/")[-1] in ("Random-v0", "Random"):
        return RandomPolicy(model, agent_id, policy_id)
    raise KeyError(f"oracle shim: policy {policy_id!r} is not available")
