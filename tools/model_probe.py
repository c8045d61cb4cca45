"""Dev probe: model.step sweep and fused rollout throughput of the generic-kernel envs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posggym_baselines_b200.envs.lbf import BatchedLBF
from posggym_baselines_b200.envs.predator_prey import BatchedPredatorPrey

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
for name, env in (("lbf", BatchedLBF()), ("pp", BatchedPredatorPrey())):
    st = env.sample_initial_state(n, seed=1)
    out = None
    def sweep(depth=10):
        global out
        cur = st
        for k in range(depth):
            out = env.step(cur, seed=1, step_idx=k, out=out, want_actions=True)
            cur = out.state
            # ping-pong would need two buffers; in-place state is fine for timing
    sweep(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sweep(20); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    b = 8 + 8 + 16 * env.obs_words + 16 + 1 + 2
    print(f"{name}: step {ms*1e3:.1f} us per launch, {n/ms/1e6:.2f} G env-steps/s, {b*n/ms/1e6:.0f} GB/s algorithmic ({b} B/env-step)")
    env.rollout(st, 0, 50, 0.95, seed=1); torch.cuda.synchronize()
    live = torch.zeros(1, dtype=torch.int64, device="cuda")
    e0.record(); env.rollout(st, 0, 50, 0.95, seed=1, live_count=live); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name}: rollout {ms:.2f} ms, {n/ms/1e6:.3f} G rollouts/s, {live.item()/ms/1e6:.2f} G env-steps/s")
