"""Dev probe: root statistics at 65 536 simulations against the reference golden for alternative wave floors of the
planner's schedule (MCTS.WAVE_FLOORS), and the time the budget takes."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy
from posggym_baselines_b200.planning.mcts import MCTS
from tests.test_pomcp_gpu import _env_model, _first_obs

G = os.path.join(ROOT, "tests", "golden")
gold = json.load(open(os.path.join(G, "planner_driving.json")))
gold_envs = json.load(open(os.path.join(G, "planner_envs.json")))
large = json.load(open(os.path.join(G, "planner_large.json")))
CANDS = {
    "A 32|64@8192": ((0, 32), (8192, 64)),
    "B 64|128@8192": ((0, 64), (8192, 128)),
    "C 32|64@2048|128@16384": ((0, 32), (2048, 64), (16384, 128)),
    "D 64|128@4096|256@32768": ((0, 64), (4096, 128), (32768, 256)),
    "E 128|256@16384": ((0, 128), (16384, 256)),
}
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
only = sys.argv[2].split(",") if len(sys.argv) > 2 else list("ABCDE")
scens = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3]
for name, floors in CANDS.items():
    if name[0] not in only:
        continue
    MCTS.WAVE_FLOORS = floors
    worst, tt = [], 0.0
    for idx, scen in enumerate(large["scenarios"]):
        if idx not in scens:
            continue
        env_id = scen["env_id"]
        model = _env_model(env_id, large["env_kwargs"][env_id])
        obs = _first_obs(scen, gold, gold_envs)
        vals, vis = [], []
        for rep in range(reps):
            cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=rep)
            p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), max_sims=large["n_sims"], node_capacity=1 << 18, log_capacity=1 << 22)
            p.reset(); p.update(None, obs)
            t0 = time.time(); p.get_action(); tt += time.time() - t0
            st = p._last_root_stats
            vals.append(np.where(st["done"] > 0, st["total"] / np.maximum(st["done"], 1), 0.0)); vis.append(st["done"] / st["done"].sum())
            p.close()
        vals, vis = np.array(vals), np.array(vis)
        ref_mean, ref_std = np.array(scen["value_mean"]), np.array(scen["value_std"])
        tol = 0.02 if env_id == "Driving-v1" else 0.003
        se = np.sqrt(ref_std**2 / large["reps"] + vals.std(0) ** 2 / reps)
        dev = (np.abs(vals.mean(0) - ref_mean) / (4 * se + tol)).max()
        dev_se = (np.abs(vals.mean(0) - ref_mean) / se).max()
        ref_vis = np.array(scen["visit_frac_mean"]); top = int(np.argmax(ref_vis))
        worst.append((round(float(dev), 2), round(float(dev_se), 1), round(float(abs(vis.mean(0)[top] - ref_vis[top])), 3)))
    print(name, "per scenario (dev / (4 SE + tol), dev / SE, top visit share diff):", worst, "ms per 65536 sims: %.1f" % (1e3 * tt / (reps * len(scens))), "values", np.round(vals.mean(0), 4).tolist(), "ref", np.round(ref_mean, 4).tolist())
