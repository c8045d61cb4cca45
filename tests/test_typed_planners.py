"""Typed-planner plugins: ``POTMMCPMetaPolicy``, ``OtherAgentMixturePolicy``,
``SearchPolicyWrapper`` and the POTMMCP differences of the host engine, against the UNMODIFIED reference classes
(imported through oracle.shims where /root/reference exists) and against hand-computed values.  CPU only: these are
host-side objects; the planner that drives them is tested on the GPU (tests/test_plugins_gpu.py)."""
import random

import pytest

from tests.potmmcp_policies import META_POLICY, BiasedPolicy, ego_policies, other_agent_types


class _TwoAgentModel:
    """The attributes the plugin classes read from a model."""

    possible_agents = ("0", "1")

    class _Space:
        n = 5

    action_spaces = {"0": _Space(), "1": _Space()}


def _returns(rng, n_ego=4, n_other=3, ties=True):
    vals = [0.0, 0.5, 1.0] if ties else None
    return {f"e{i}": {f"o{j}": (rng.choice(vals) if ties else rng.uniform(-1, 1)) for j in range(n_other)}
            for i in range(n_ego)}


def test_meta_policy_builders_hand_values():
    from posggym_baselines_b200.planning import POTMMCPMetaPolicy as MP

    ret = {"a": {"x": 1.0, "y": 0.0}, "b": {"x": 1.0, "y": 2.0}, "c": {"x": 0.0, "y": -1.0}}
    assert MP.get_uniform_meta_policy(ret) == {"x": {"a": 1 / 3, "b": 1 / 3, "c": 1 / 3}, "y": {"a": 1 / 3, "b": 1 / 3, "c": 1 / 3}}
    assert MP.get_greedy_meta_policy(ret) == {"x": {"a": 0.5, "b": 0.5, "c": 0.0}, "y": {"a": 0.0, "b": 1.0, "c": 0.0}}
    soft = MP.get_softmax_meta_policy(ret, 0.5)
    import math

    z = math.exp(0) + math.exp(4) + math.exp(-2)
    assert soft["y"] == pytest.approx({"a": 1 / z, "b": math.exp(4) / z, "c": math.exp(-2) / z})
    assert all(abs(sum(d.values()) - 1) < 1e-12 for d in soft.values())


def test_meta_policy_builders_match_reference(oracle_env):
    if not oracle_env["have_reference"]:
        pytest.skip("needs the reference tree (build container)")
    from posggym_baselines.planning import POTMMCPMetaPolicy as Ref
    from posggym_baselines_b200.planning import POTMMCPMetaPolicy as MP

    rng = random.Random(0)
    for k in range(50):
        ret = _returns(rng, ties=k % 2 == 0)
        assert MP.get_uniform_meta_policy(ret) == Ref.get_uniform_meta_policy(ret)
        assert MP.get_greedy_meta_policy(ret) == Ref.get_greedy_meta_policy(ret)
        for temp in (0.25, 1.0, 4.0):
            assert MP.get_softmax_meta_policy(ret, temp) == Ref.get_softmax_meta_policy(ret, temp)   # bit-equal floats


def test_meta_policy_refuses_scalar_calls_and_validates():
    from posggym_baselines_b200.planning import POTMMCPMetaPolicy

    m = _TwoAgentModel()
    mp = POTMMCPMetaPolicy(m, "0", ego_policies(m), META_POLICY)
    st = mp.get_initial_state()
    assert set(st) == {"fast", "careful"} and mp.other_agent_id == "1" and mp.policy_id == "POTMMCPMetaPolicy"
    for call in (mp.sample_action, mp.get_pi, mp.get_value):
        with pytest.raises(NotImplementedError):
            call(st)
    with pytest.raises(AssertionError):
        POTMMCPMetaPolicy(m, "0", ego_policies(m), {"pusher": {"fast": 0.5, "careful": 0.4}})
    with pytest.raises(AssertionError):
        POTMMCPMetaPolicy(m, "0", ego_policies(m), {"pusher": {"unknown": 1.0}})


def test_expected_action_probs_hand_value():
    from posggym_baselines_b200.planning import POTMMCPMetaPolicy

    m = _TwoAgentModel()
    pol = ego_policies(m)
    mp = POTMMCPMetaPolicy(m, "0", pol, META_POLICY)
    st = mp.get_next_state(None, "obs", mp.get_initial_state())
    assert st == {"fast": {"k": 1}, "careful": {"k": 1}}
    # uniform over the two types: P(fast) = (0.25 + 0.9) / 2
    w_fast = (0.25 + 0.9) / 2
    want = {a: w_fast * pol["fast"].get_pi(st["fast"]).probs[a] + (1 - w_fast) * pol["careful"].get_pi(st["careful"]).probs[a]
            for a in range(5)}
    got = mp.get_expected_action_probs(None, st)
    assert got == pytest.approx(want) and abs(sum(got.values()) - 1) < 1e-12
    only = mp.get_expected_action_probs({"idler": 1.0}, st)
    assert only[1] > got[1]                       # against an idler the meta-policy mostly drives fast


def test_plugin_classes_match_reference_call_for_call(oracle_env):
    """Same seeds -> same draws: the meta-policy's sampled policies and expected distributions, the mixture policy's
    types / states / actions and the wrapper's padded distributions are those of the reference classes."""
    if not oracle_env["have_reference"]:
        pytest.skip("needs the reference tree (build container)")
    import posggym_baselines.planning as ref
    import posggym_baselines_b200.planning as new

    m = _TwoAgentModel()
    out = []
    for mod in (ref, new):
        random.seed(5)
        mp = mod.POTMMCPMetaPolicy(m, "0", ego_policies(m), META_POLICY)
        mix = mod.OtherAgentMixturePolicy(m, "1", other_agent_types(m))
        trace = []
        sp_state = mp.get_initial_state()
        for ep in range(20):
            o_state = mix.sample_initial_state()
            for t in range(6):
                a = mix.sample_action(o_state)
                pi = mp.sample_policy({"1": o_state})
                assert isinstance(pi, mod.SearchPolicyWrapper)
                ego_state = sp_state[pi.policy_id]
                trace.append((o_state["policy_id"], a, pi.policy_id, pi.sample_action(ego_state), tuple(pi.get_pi(ego_state).items()),
                              tuple(mix.get_pi(o_state).items()), tuple(mp.get_expected_action_probs(None, sp_state).items()),
                              tuple(mp.get_expected_action_probs({"pusher": 0.3, "idler": 0.7}, sp_state).items())))
                o_state = mix.get_next_state(a, ("obs", t), o_state)
                sp_state = mp.get_next_state(a, ("obs", t), sp_state)
        hist = [(None, "o0"), (1, "o1"), (2, "o2")]
        trace.append(mix.get_state_from_history(mix.sample_initial_state(), hist))
        out.append(trace)
    assert out[0] == out[1]


def test_wrapper_pads_sparse_distributions():
    from posggym_baselines_b200.planning import OtherAgentMixturePolicy, SearchPolicyWrapper

    class Sparse(BiasedPolicy):
        def get_pi(self, state):
            return type("D", (), {"probs": {1: 0.75, 3: 0.25}})()

    m = _TwoAgentModel()
    w = SearchPolicyWrapper(Sparse(m, "0", "sparse", 1, 1.0, value=0.5))
    assert w.get_pi({"k": 0}) == {1: 0.75, 3: 0.25, 0: 0.0, 2: 0.0, 4: 0.0} and w.get_value({"k": 0}) == 0.5
    assert w.policy_id == "sparse" and w.agent_id == "0"
    mix = OtherAgentMixturePolicy(m, "1", {"s": Sparse(m, "1", "s", 1, 1.0)})
    st = mix.sample_initial_state()
    assert st == {"policy_id": "s", "policy_state": {"k": 0}} and set(mix.get_pi(st)) == set(range(5))


def test_potmmcp_engine_differences():
    """The four hooks of planning/potmmcp.py on bare tree nodes (no device needed): root prior = expected
    distribution, simulation prior = the sampled policy's, leaf state = that policy's entry, and a revisit folds the
    sampled policy's distribution into the PARENT node's prior with the parent's visit count (potmmcp.py:244-252)."""
    from posggym_baselines_b200.planning import POTMMCPMetaPolicy
    from posggym_baselines_b200.planning.host_search import HPS, _Belief, _ObsNode
    from posggym_baselines_b200.planning.potmmcp import _PotmmcpEngine

    m = _TwoAgentModel()
    mp = POTMMCPMetaPolicy(m, "0", ego_policies(m), META_POLICY)
    eng = object.__new__(_PotmmcpEngine)
    eng.search_policy = mp
    st = mp.get_next_state(None, "o", mp.get_initial_state())
    assert eng._node_prior(st, None) == mp.get_expected_action_probs(None, st)
    random.seed(1)
    sim = eng._sim_policy(HPS("s", (), {"1": {"policy_id": "idler", "policy_state": {"k": 0}}}, 1))
    assert sim.policy_id in ("fast", "careful")
    assert eng._node_prior(st, sim) == sim.get_pi(st[sim.policy_id])
    node = _ObsNode(None, "o", 1, _Belief(random.Random(0)), dict(eng._node_prior(st, None)), st, visits=4)
    assert eng._leaf_state(node, sim) == st[sim.policy_id]
    child = _ObsNode(None, "o2", 2, _Belief(random.Random(0)), {a: 0.2 for a in range(5)}, st, visits=1)
    before, child_before = dict(node.action_probs), dict(child.action_probs)
    eng._revisit(node, child, sim)
    pi = sim.get_pi(st[sim.policy_id])
    assert child.visits == 2 and child.action_probs == child_before
    assert node.action_probs == pytest.approx({a: before[a] + (pi[a] - before[a]) / 4 for a in range(5)})
    # 'idler' draws fast with probability 0.9
    random.seed(2)
    ids = [eng._sim_policy(HPS("s", (), {"1": {"policy_id": "idler", "policy_state": {}}}, 1)).policy_id for _ in range(400)]
    assert 0.84 < ids.count("fast") / 400 < 0.96


def test_typed_planners_need_history_particles():
    from posggym_baselines_b200.planning import IPOMCP, MCTS, POTMMCP, MCTSConfig

    assert issubclass(IPOMCP, MCTS) and issubclass(POTMMCP, MCTS)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.25, truncated=False, state_belief_only=True)
    with pytest.raises(ValueError):
        POTMMCP(_TwoAgentModel(), "0", cfg, {}, None)


def test_scalar_codec_memoises_both_directions():
    """envs/base.py ``ScalarCodec``: the scalar facade step's packed-word <-> value conversions are looked up, not
    recomputed; a state first seen as device output is fed back with exactly the device's words."""
    import numpy as np

    from posggym_baselines_b200.envs.base import ScalarCodec

    class FakeBatched:
        calls = {"pack": 0, "unpack_state": 0, "unpack_obs": 0}

        def pack_state(self, states):
            self.calls["pack"] += 1
            return np.array([[s[0] * 10, s[1] * 10] for s in states], dtype=np.int32)

        def unpack_state(self, packed):
            self.calls["unpack_state"] += 1
            return [(int(r[0]) // 10, int(r[1]) // 10) for r in np.asarray(packed)]

        def unpack_obs(self, packed):
            self.calls["unpack_obs"] += 1
            return [tuple(("o", int(v)) for v in row) for row in np.asarray(packed)]

    b = FakeBatched()
    codec = ScalarCodec(b)
    row = codec.pack((1, 2))
    assert row.tolist() == [[10, 20]] and codec.pack((1, 2)) is row and b.calls["pack"] == 1
    dev = np.array([[31, 47]], dtype=np.int32)                      # a device word that pack_state would not produce
    st = codec.state(dev)
    assert st == (3, 4) and codec.state(dev.copy()) is st and b.calls["unpack_state"] == 1
    assert codec.pack(st).tolist() == [[31, 47]] and b.calls["pack"] == 1     # fed back as the device wrote it
    ob = codec.obs(np.array([7, 9], dtype=np.int64))
    assert ob == (("o", 7), ("o", 9)) and codec.obs(np.array([7, 9], dtype=np.int64)) is ob and b.calls["unpack_obs"] == 1
    codec.LIMIT = 2                                                  # bounded: the memo is dropped when full
    for k in range(5):
        codec.pack((k + 10, k))
    assert len(codec._packed) <= 2
