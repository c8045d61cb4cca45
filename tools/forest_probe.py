"""Dev probe: POMCP with K trees per GPU (root parallelisation inside the GPU): simulations per 0.1 s decision and
update time for several K."""
import os, sys, random, json, time, gc
if os.environ.get("NOGC"): gc.disable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from posggym_baselines_b200.envs.env import make
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy
from posggym_baselines_b200.planning.mcts import MCTS
if os.environ.get("UPDATE_THREADS"): MCTS.UPDATE_THREADS = int(os.environ["UPDATE_THREADS"])

for K in [int(a) for a in sys.argv[1:]] or [1, 8, 32]:
    env = make("Driving-v1")
    model = env.model
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=3)
    t0 = time.time()
    p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), trees_per_gpu=K)
    rng = random.Random(3)
    obs, _ = env.reset(seed=3)
    p.reset()
    rows = []
    for d in range(9):
        a = p.step(obs["0"])
        if p.root.is_absorbing:
            break
        st, info = p.step_statistics, p.search_info
        tr = sorted(getattr(p, "_update_trace", []), key=lambda r: -r[3])[:3]
        if tr and st["update_time"] > 0.08:
            print("   slow update: total %.3f; slowest trees (idx, thread, start, dur):" % st["update_time"], [(a, b, round(c, 4), round(d, 4), e) for a, b, c, d, e in tr],
                  "sum of durations %.3f" % sum(r[3] for r in p._update_trace))
        rows.append((info["num_sims"], info["sims_per_tree"], round(st["search_time"], 4), round(st["update_time"], 4),
                     info["waves"], info["last_wave"], round(info["arena_fill"], 4)))
        obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        if done:
            break
    print(K, "setup+episode %.1f s" % (time.time() - t0))
    for r in rows:
        print("   ", r)
    p.close()
