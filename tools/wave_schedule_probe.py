"""Dev probe: root statistics after 65 536 simulations for several wave schedules (wave = growth x root
visits) against the unmodified reference planner's golden (tests/golden/planner_large.json), and the
simulations a 0.1 s decision gets under each schedule."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from posggym_baselines_b200.envs.env import Env, make  # noqa: E402
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
large = json.load(open(os.path.join(G, "planner_large.json")))
drv = json.load(open(os.path.join(G, "planner_driving.json")))
envs = json.load(open(os.path.join(G, "planner_envs.json")))


def f2o(f):
    return (tuple(f[:15]), f[15], (f[16], f[17]), (f[18], f[19]), f[20], f[21])


growths = [float(x) for x in sys.argv[1:]] or [1.0, 0.25, 1 / 16, 1 / 64]
for scen in large["scenarios"]:
    env_id = scen["env_id"]
    model = make(env_id, **large["env_kwargs"][env_id]).model
    if env_id == "Driving-v1":
        obs = f2o(next(r for r in drv["root_stats"] if r["env_seed"] == scen["env_seed"])["ego_obs"])
    else:
        obs = tuple(next(r for r in envs[env_id]["root_stats"] if r["env_seed"] == scen["env_seed"])["ego_obs"])
    print(env_id, scen["env_seed"], "reference", np.round(scen["value_mean"], 4), np.round(scen["visit_frac_mean"], 3))
    for g in growths:
        vals, vis, dt, waves = [], [], [], []
        for rep in range(6):
            cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=rep)
            p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), max_sims=large["n_sims"], wave_growth=g,
                      node_capacity=1 << 18, log_capacity=1 << 22)
            p.reset(); p.update(None, obs)
            t0 = time.time(); p.get_action(); dt.append(time.time() - t0)
            st = p._last_root_stats
            vals.append(np.where(st["done"] > 0, st["total"] / np.maximum(st["done"], 1), 0.0))
            vis.append(st["done"] / st["done"].sum()); waves.append(p.search_info["waves"])
            p.close()
        print(f"  growth {g:7.4f}: values", np.round(np.mean(vals, 0), 4), "visits", np.round(np.mean(vis, 0), 3),
              f"waves {np.mean(waves):.0f} time {1e3 * np.mean(dt):.1f} ms")
# simulations per 0.1 s decision in an episode under each schedule
model = make("Driving-v1").model
for g in growths:
    env = Env(model)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=0)
    p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), wave_growth=g)
    obs, _ = env.reset(seed=100); p.reset()
    sims, upd, created = [], [], []
    for d in range(8):
        a = p.step(obs["0"])
        if d > 0:
            sims.append(p.step_statistics["num_sims"]); upd.append(p.step_statistics["update_time"])
            created.append(p.search_info["nodes_created"])
        obs, _, _, _, done, _ = env.step({"0": a, "1": 0})
        if done:
            break
    print(f"growth {g:7.4f}: sims/decision {np.mean(sims):.3e} nodes created {np.mean(created):.3e} update {1e3*np.mean(upd):.1f} ms "
          f"overflow {p.search_info['overflow_share']:.3f}")
    p.close()
