"""``posggym.agents``-style policy objects for the typed-planner tests (duck-typed: ``model``, ``agent_id``,
``policy_id``, ``get_initial_state`` / ``get_next_state`` / ``sample_action`` / ``get_pi(...).probs`` /
``get_value`` / ``close``).  Shared by tests/golden/make_potmmcp_golden.py, which drives the UNMODIFIED reference
POTMMCP with them, and by the tests of this repo's POTMMCP."""
import random


class _Dist:
    def __init__(self, probs):
        self.probs = probs


class BiasedPolicy:
    """Prefers one action, less so the longer the episode; state = number of steps seen."""

    def __init__(self, model, agent_id, policy_id, favourite, weight, value=None):
        self.model, self.agent_id, self.policy_id = model, agent_id, policy_id
        self.n = model.action_spaces[agent_id].n
        self.favourite, self.weight, self.value = favourite, weight, value

    def get_initial_state(self):
        return {"k": 0}

    def get_next_state(self, action, obs, state):
        return {"k": state["k"] + 1}

    def get_pi(self, state):
        w = [1.0] * self.n
        w[self.favourite] += self.weight / (1.0 + 0.1 * state["k"])
        z = sum(w)
        return _Dist({a: w[a] / z for a in range(self.n)})

    def sample_action(self, state):
        pi = self.get_pi(state).probs
        return random.choices(list(pi), weights=list(pi.values()), k=1)[0]

    def get_value(self, state):
        if self.value is None:
            raise NotImplementedError
        return self.value

    def close(self):
        pass


def other_agent_types(model, agent_id="1"):
    """The two types the other driver may be: keeps accelerating / mostly idles."""
    return {"pusher": BiasedPolicy(model, agent_id, "pusher", 1, 6.0),
            "idler": BiasedPolicy(model, agent_id, "idler", 0, 6.0)}


def ego_policies(model, agent_id="0"):
    return {"fast": BiasedPolicy(model, agent_id, "fast", 1, 4.0),
            "careful": BiasedPolicy(model, agent_id, "careful", 2, 2.0)}


META_POLICY = {"pusher": {"fast": 0.25, "careful": 0.75}, "idler": {"fast": 0.9, "careful": 0.1}}


def plugin_policies_for(model, levels, wrapper):
    """Search policies per nesting level and agent for the general I-NTMCP tests: level l, agent i prefers ACCELERATE
    (agent 0) / DECELERATE (agent 1), more strongly at higher levels."""
    return {lvl: {i: wrapper(BiasedPolicy(model, i, f"L{lvl}_{i}", 1 if i == "0" else 2, 2.0 + 2.0 * lvl))
                  for i in model.possible_agents} for lvl in range(levels + 1)}
