"""Dev probe (torchrun, N ranks): POMCP with root_parallel=True AND trees_per_gpu=K -- the K local trees are summed on
the rank, then the ranks' sums merged with the one allreduce; every rank must take the same action."""
import os, sys, random
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from posggym_baselines_b200.envs.env import Env
from posggym_baselines_b200.envs.predator_prey import PredatorPreyModel
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

model = PredatorPreyModel(device=torch.device("cuda", local))
env = Env(model)
cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=7)
p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), root_parallel=True, trees_per_gpu=int(sys.argv[1]) if len(sys.argv) > 1 else 4)
rng = random.Random(7)
obs, _ = env.reset(seed=7)
p.reset()
acts, sims = [], []
for d in range(6):
    a = p.step(obs["0"]); acts.append(a); sims.append(p.search_info["num_sims"])
    obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
    if done:
        break
t = torch.tensor(acts + [-1] * (6 - len(acts)), device="cuda")
ref = t.clone(); dist.broadcast(ref, 0)
same = torch.tensor([int(torch.equal(ref, t))], device="cuda"); dist.all_reduce(same, op=dist.ReduceOp.MIN)
tot = torch.tensor([float(sum(sims))], device="cuda"); dist.all_reduce(tot)
if rank == 0:
    print("ranks agree on every action:", bool(same.item()), "actions", acts, "simulations per decision over all ranks: %.3e" % (tot.item() / len(sims)),
          "merged root visits", p._merged_root[0].sum())
dist.destroy_process_group()
