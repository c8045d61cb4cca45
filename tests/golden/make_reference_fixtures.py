"""Distil the in-tree fixtures of /root/reference that pin the hot path's model
semantics into one small JSON that travels with the repo
(the reference tree itself does not exist on the GPU box).  Run in the build container:
`python tests/golden/make_reference_fixtures.py`."""
from __future__ import annotations

import glob
import json
import os
import pickle
from collections import Counter
from fractions import Fraction

import numpy as np
import pandas as pd
import torch

REF = "/root/reference/baseline_exps/env_data"
HERE = os.path.dirname(os.path.abspath(__file__))


class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, s):
        if isinstance(s, dict):
            self.__dict__.update(s)


class _U(pickle.Unpickler):
    def find_class(self, mod, name):
        if mod.startswith(("gymnasium", "posggym")):
            return type(name, (_Stub,), {})
        return super().find_class(mod, name)


class _M:
    Unpickler = _U
    load = pickle.load
    __name__ = "M"


def load_ckpt(f):
    return torch.load(f, map_location="cpu", weights_only=False, pickle_module=_M)


def checkpoint_fixtures(env):
    """obs width, #actions, and which flattened one-hot inputs were EVER active during
    training (Adam's exp_avg_sq of the first layer is exactly 0 for never-active inputs)."""
    active = None
    for f in sorted(glob.glob(f"{REF}/{env}/br_models/*.pt")):
        ck = load_ckpt(f)
        v = ck["optimizer"]["state"][0]["exp_avg_sq"].numpy()
        a = (v != 0).any(0)
        active = a if active is None else (active | a)
        obs_dim = int(ck["model"]["trunk.0.weight"].shape[1])
        n_act = int(ck["model"]["actor.2.weight"].shape[0])
        kwargs = ck["config"]["env_kwargs"]
    return {"obs_dim": obs_dim, "num_actions": n_act, "env_kwargs": kwargs,
            "active_inputs": "".join("1" if x else "0" for x in active)}


def driving_reward_fixtures():
    df = pd.read_csv(f"{REF}/Driving-v1/br_results.csv")
    r = df["return"].values
    succ = df[np.abs(r - 1) < 1e-6]
    resid = np.round((succ["return"].values - 1) * 2**32).astype(int)
    resid_hist = {int(k): int(v) for k, v in sorted(Counter(resid).items())}
    min_len = {int(k): int(succ["len"].values[resid == k].min()) for k in resid_hist}
    fails = df[np.abs(r - 1) >= 1e-6]
    denoms = Counter()
    for x in fails["return"].values:
        base = x + 1 if x < 0 else x
        denoms[Fraction(float(base)).limit_denominator(60).denominator] += 1
    return {"n_episodes": int(len(df)), "n_success": int(len(succ)),
            "success_residual_hist_2p32": resid_hist, "success_min_len_by_residual": min_len,
            "failure_return_denominators": {int(k): int(v) for k, v in sorted(denoms.items())},
            "max_len": int(df["len"].max())}


def random_play(env):
    df = pd.read_csv(f"{REF}/{env}/div_results.csv")
    rows = df[(df.policy_id == "Random-v0") & (df.co_team_id == "(Random)")]
    out = []
    for _, x in rows.iterrows():
        out.append({k: (x[k].item() if hasattr(x[k], "item") else x[k]) for k in
                    ["agent_id", "num_episodes", "episode_reward_mean", "episode_reward_std",
                     "num_LOSS", "num_WIN", "num_NA", "num_DRAW"]})
    solo = df[(df.policy_id == "Random-v0") & (df.num_LOSS == 0)]
    return {"random_vs_random": out,
            "random_uncrashed_reward_mean": sorted(set(round(float(v), 9) for v in solo.episode_reward_mean))}


def lattice(env):
    df = pd.read_csv(f"{REF}/{env}/br_results.csv")
    vals = Counter(Fraction(float(x)).limit_denominator(1000) for x in df["return"].values)
    top = sorted(vals.items(), key=lambda kv: -kv[1])[:40]
    return {"max_len": int(df["len"].max()), "top_returns": [[str(k), int(v)] for k, v in top]}


if __name__ == "__main__":
    out = {
        "source": "derived from /root/reference/baseline_exps/env_data by tests/golden/make_reference_fixtures.py",
        "Driving-v1": {**checkpoint_fixtures("Driving-v1"), "rewards": driving_reward_fixtures(),
                       **random_play("Driving-v1")},
        "PredatorPrey-v0": {**checkpoint_fixtures("PredatorPrey-v0"), **random_play("PredatorPrey-v0"),
                            "lattice": lattice("PredatorPrey-v0")},
        "LevelBasedForaging-v3": {**checkpoint_fixtures("LevelBasedForaging-v3"),
                                  **random_play("LevelBasedForaging-v3"),
                                  "lattice": lattice("LevelBasedForaging-v3")},
    }
    with open(os.path.join(HERE, "reference_fixtures.json"), "w") as f:
        json.dump(out, f, indent=1, default=str)
    print(json.dumps(out, indent=1, default=str)[:6000])
