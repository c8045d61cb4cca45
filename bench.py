#!/usr/bin/env python
"""bench.py -- BASELINE.json metric on configs[1]: batched Driving-v1 model.step
random-action rollout sweep (16M particles per GPU, depth 50), env-steps/s.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

One "step" = one depth-50 sweep over the particle batch through the step-API kernel
(50 launches; every launch reads the packed state and writes state, packed obs, float64
rewards, flags and the sampled joint action: 51 B per env-step, HBM-bound).  Prints ONE
JSON line (rank 0).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "driving_v1_model_env_steps_per_s"
UNIT = "env-steps/s"
BYTES_PER_ENV_STEP = 51  # 8 B state in; 8 state + 16 obs + 16 rewards + 1 flags + 2 actions out
DEPTH = 50


def workload_config(n: int, seed: int = 2026) -> dict:
    """``config`` of the JSON line -- identical for this repo's arm and the reference arm."""
    return {"workload": f"Driving-v1 14x14RoundAbout, 2 agents, obs_dim (3,1,1): random-action model.step sweep, "
                        f"{n} particles per GPU x depth {DEPTH} (BASELINE.json configs[1])",
            "particles_per_gpu": n, "depth": DEPTH, "seed": seed,
            "l2": f"per-launch working set {n * BYTES_PER_ENV_STEP / 1e6:.0f} MB >> 126 MB L2 (inputs larger than L2)",
            "counted": "env-steps executed on not-yet-all-done particles"}


def load_counters() -> dict:
    """Per-kernel counters distilled from the committed ncu captures (profiles/kernel_counters.json)."""
    p = os.path.join(ROOT, "profiles", "kernel_counters.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                mx = max(mx, float(parts[1]))
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(parts[0]))
                    for nm, v in zip(names, parts[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """CPU reference arm: the oracle port of the posggym Driving-v1 model (pure Python,
    like the reference), one process per host core, same random-action sweep."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_bench

    cores = os.cpu_count() or 1
    budget = max(1.0, min(8.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(min(args.warmup, 1)):
        cpu_bench.driving_sweep(budget_s=1.0, depth=DEPTH, procs=cores, seed=1)
    vals, secs = [], 0.0
    for k in range(args.steps):
        r = cpu_bench.driving_sweep(budget_s=budget, depth=DEPTH, procs=cores, seed=100 + k)
        vals.append(r["env_steps_per_s"])
        secs += r["seconds"]
    v = sum(vals) / len(vals)
    sample = (f"bounded sample of the workload: {budget:.1f} s of Driving-v1 random-action depth-{DEPTH} episodes per step "
              f"on each of {cores} processes (Python oracle port of the model; the reference's model is pure Python)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args.particles),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


FOREST_TREES = 128


def bench_pomcp(dev, rank, world, cpu: bool, decisions: int = 8):
    """README configuration (README.md:52-88): POMCP on Driving-v1 vs a uniform-random other
    agent, discount 0.95, c 1.414, search_time_limit 0.1 s, UCB.  Simulations per decision on
    the GPU (one search tree per rank) next to the restated reference planner on one host core
    (the reference planner is single-threaded)."""
    import random

    import torch
    import torch.distributed as dist

    from posggym_baselines_b200.envs.driving import DrivingModel
    from posggym_baselines_b200.envs.env import Env
    from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

    model = DrivingModel(device=dev)
    env = Env(model)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=rank)
    def episode(trees_per_gpu):
        planner = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), trees_per_gpu=trees_per_gpu)
        rng = random.Random(rank)
        obs, _ = env.reset(seed=100 + rank)
        planner.reset()
        sims, times, upd, infos = [], [], [], []
        for d in range(decisions + 1):
            a = planner.step(obs["0"])
            if d > 0 and planner.step_statistics["num_sims"] == planner.step_statistics["num_sims"]:  # skip warm-up / NaN
                sims.append(planner.step_statistics["num_sims"])
                times.append(planner.step_statistics["search_time"])
                upd.append(planner.step_statistics["update_time"])
                infos.append(dict(planner.search_info))
            obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
            if done:
                obs, _ = env.reset()
                planner.reset()
        cap = planner._cap_nodes
        planner.close()
        return sims, times, upd, infos, cap

    sims, times, upd, infos, cap_nodes = episode(1)
    tot = torch.tensor([sum(sims), len(sims)], dtype=torch.float64, device=dev)
    tmax = torch.tensor([sum(times) / max(1, len(times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    out = {"config": "Driving-v1, discount 0.95, c 1.414, search_time_limit 0.1 s, ucb, RandomSearchPolicy",
           "sims_per_decision_per_gpu": tot[0].item() / max(1.0, tot[1].item()),
           "sims_per_decision_all_gpus": world * tot[0].item() / max(1.0, tot[1].item()),
           "sims_per_s_all_gpus": world * tot[0].item() / max(1.0, tot[1].item()) / max(1e-9, tmax.item()),
           "search_time_s": tmax.item(), "trees": world,
           # planner.update (re-root + belief + reinvigoration) runs outside the search budget, like the reference's
           # update_time; the next four say how much of the search still grew the tree (rank 0's tree)
           "update_time_s": sum(upd) / max(1, len(upd)),
           "overflow_share": sum(i["overflow_share"] for i in infos) / max(1, len(infos)),
           "nodes_created_per_decision": sum(i["nodes_created"] for i in infos) / max(1, len(infos)),
           "arena_fill_max": max([i["arena_fill"] for i in infos] or [0.0]),
           "waves_per_decision": sum(i["waves"] for i in infos) / max(1, len(infos)),
           "node_capacity": cap_nodes}
    # Root parallelisation INSIDE the GPU (trees_per_gpu): FOREST_TREES independent trees, one CTA / SM each, searched
    # by one launch per chunk and merged at the root -- the same merge as the cross-GPU allreduce.  Every tree is
    # the validated single-tree search (tests/test_pomcp_gpu.py::test_forest_*), so a decision sees that many
    # times the simulations of one tree; the update (re-root + belief per tree) is still sequential over the trees.
    try:
        fs, ft, fu, fi, fcap = episode(FOREST_TREES)
        ftot = torch.tensor([sum(fs), len(fs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ftot)
        out["root_parallel_per_gpu"] = {
            "trees_per_gpu": FOREST_TREES,
            "sims_per_decision_per_gpu": ftot[0].item() / max(1.0, ftot[1].item()),
            "sims_per_decision_all_gpus": world * ftot[0].item() / max(1.0, ftot[1].item()),
            "sims_per_tree_per_decision": sum(i["sims_per_tree"] for i in fi) / max(1, len(fi)),
            "search_time_s": sum(ft) / max(1, len(ft)),
            # (the first decisions of the first forest in a process pay one-time costs -- lazily loaded kernels, the
            # stream-ordered allocator's pool growing on eight streams: the median is the steady state)
            "update_time_s": sorted(fu)[len(fu) // 2] if fu else 0.0, "update_time_mean_s": sum(fu) / max(1, len(fu)),
            "overflow_share": sum(i["overflow_share"] for i in fi) / max(1, len(fi)),
            "arena_fill_max": max([i["arena_fill"] for i in fi] or [0.0]), "node_capacity_per_tree": fcap}
    except Exception as ex:
        out["root_parallel_per_gpu"] = {"error": repr(ex)}
    # BASELINE.json configs[2]: root-parallel POMCP on PredatorPrey-v0 -- every rank grows its own
    # tree from the same root belief; root visit counts and value sums merged by ONE allreduce
    # per decision (NCCL); every rank then takes the same action.
    try:
        from posggym_baselines_b200.envs.predator_prey import PredatorPreyModel

        pp_model = PredatorPreyModel(device=dev)
        pp_env = Env(pp_model)
        pp_cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=7)
        pp_planner = POMCP(pp_model, "0", pp_cfg, RandomSearchPolicy(pp_model, "0"), root_parallel=True)
        pp_rng = random.Random(7)
        pp_obs, _ = pp_env.reset(seed=7)   # identical env trajectory on every rank
        pp_planner.reset()
        pp_sims, pp_ar, pp_actions = [], [], []
        for d in range(decisions + 1):
            a = pp_planner.step(pp_obs["0"])
            if d > 0 and pp_planner.step_statistics["num_sims"] == pp_planner.step_statistics["num_sims"]:
                pp_sims.append(pp_planner.step_statistics["num_sims"])
                pp_ar.append(pp_planner.allreduce_time)
            pp_actions.append(a)
            pp_obs, _, _, _, done, _ = pp_env.step({"0": a, "1": pp_rng.randrange(5)})
            if done:
                break
        pp_planner.close()
        tot_pp = torch.tensor([sum(pp_sims), len(pp_sims)], dtype=torch.float64, device=dev)
        act_t = torch.tensor(pp_actions + [0] * (decisions + 1 - len(pp_actions)), dtype=torch.int64, device=dev)
        agree = True
        if world > 1:
            dist.all_reduce(tot_pp)
            ref = act_t.clone()
            dist.broadcast(ref, 0)
            agree_t = torch.tensor([int(torch.equal(ref, act_t))], device=dev)
            dist.all_reduce(agree_t, op=dist.ReduceOp.MIN)
            agree = bool(agree_t.item())
        out["predator_prey_root_parallel"] = {
            "config": "PredatorPrey-v0 10x10, 2 predators, 3 prey, same MCTSConfig, one tree per GPU",
            "total_sims_per_decision": tot_pp[0].item() / max(1.0, tot_pp[1].item() / world),
            "allreduce_us_mean": 1e6 * sum(pp_ar) / max(1, len(pp_ar)), "allreduce_bytes": 80,
            "ranks_agree_on_actions": agree}
    except Exception as ex:
        out["predator_prey_root_parallel"] = {"error": repr(ex)}
    # BASELINE.json configs[3]: I-NTMCP nesting level 1 on Driving-v1 (tests/planning/test_intmcp.py:34-48:
    # search_time_limit 0.2 s, UCB, state_belief_only=False); two device trees per rank.
    try:
        import json as _json
        import math as _math

        from posggym_baselines_b200.planning import INTMCP

        icfg = MCTSConfig(discount=0.95, search_time_limit=0.2, c=_math.sqrt(2), truncated=False,
                          action_selection="ucb", state_belief_only=False, seed=rank)
        ipl = INTMCP.initialize(model, "0", icfg, nesting_level=1, search_policies=None)
        irng = random.Random(50 + rank)
        iobs, _ = env.reset(seed=200 + rank)
        ipl.reset()
        isims, itimes = [], []
        for d in range(6):
            a = ipl.step(iobs["0"])
            if d > 0 and ipl.step_statistics["num_sims"] == ipl.step_statistics["num_sims"] and not ipl._absorbing:
                isims.append(ipl.step_statistics["num_sims"])
                itimes.append(ipl.step_statistics["search_time"])
            iobs, _, _, _, done, _ = env.step({"0": a, "1": irng.randrange(5)})
            if done:
                break
        ipl.close()
        ref_sims = None
        gp = os.path.join(ROOT, "tests", "golden", "intmcp_driving.json")
        if os.path.exists(gp):
            ref_sims = _json.load(open(gp))["episodes"]["mean_sims_per_decision"]
        out["intmcp_level1"] = {
            "config": "Driving-v1, nesting_level 1, discount 0.95, search_time_limit 0.2 s, ucb",
            "sims_per_decision_per_gpu": sum(isims) / max(1, len(isims)),
            "search_time_s": sum(itimes) / max(1, len(itimes)),
            "note": "level-0 simulations per decision are budgeted (node_capacity / (step_limit + 2)): that tree is append-only",
            "reference_sims_per_decision_build_container": ref_sims,
            "reference_source": "unmodified reference INTMCP on the CPU oracle, 1 core, tests/golden/make_intmcp_golden.py"}
    except Exception as ex:
        out["intmcp_level1"] = {"error": repr(ex)}
    if rank == 0:
        out["plugin_path"] = bench_plugin_path(model, env)
    if cpu:
        from oracle import shims

        shims.install()
        from oracle.driving import DrivingModel as CpuModel
        from oracle.pomcp import OraclePOMCP

        m = CpuModel()
        m.seed(0)
        m.spec = type("S", (), {"max_episode_steps": 50})()
        p = OraclePOMCP(m, "0", 0.95, 0.1, 1.414, seed=0)
        s = m.sample_initial_state()
        o = m.sample_initial_obs(s)
        p.reset()
        c_sims = []
        rng = random.Random(rank)
        for d in range(12):
            a = p.step(o["0"])
            if p.step_statistics["num_sims"] == p.step_statistics["num_sims"]:
                c_sims.append(p.step_statistics["num_sims"])
            ts = m.step(s, {"0": a, "1": rng.randrange(5)})
            s, o = ts.state, ts.observations
            if ts.all_done:
                break
        out["cpu_reference_sims_per_decision"] = sum(c_sims) / max(1, len(c_sims))
        out["cpu_reference"] = "oracle/pomcp.py (restated reference planner, pure Python, 1 core), same config"
        out["ratio"] = out["sims_per_decision_per_gpu"] / max(1e-9, out["cpu_reference_sims_per_decision"])
    return out


class _TiltedPolicy:
    """A ``posggym.agents``-style policy object for the plugin-path section: prefers one action, has a value."""

    def __init__(self, model, agent_id, policy_id, favourite):
        self.model, self.agent_id, self.policy_id, self.favourite = model, agent_id, policy_id, favourite
        self.n = model.action_spaces[agent_id].n

    def get_initial_state(self):
        return {"k": 0}

    def get_next_state(self, action, obs, state):
        return {"k": state["k"] + 1}

    def get_pi(self, state):
        w = [1.0] * self.n
        w[self.favourite] += 4.0
        return type("D", (), {"probs": {a: w[a] / sum(w) for a in range(self.n)}})()

    def sample_action(self, state):
        pi = self.get_pi(state).probs
        return random.choices(list(pi), weights=list(pi.values()), k=1)[0]

    def get_value(self, state):
        return 0.05

    def close(self):
        pass


def bench_plugin_path(model, env, decisions: int = 4):
    """Planners whose policies are Python plugins (typed planners; general I-NTMCP): the tree descent
    and the plugin calls run on the host (planning/host_search.py, potmmcp.py, host_intmcp.py), the model on the
    device -- host speed by construction, reported so that nobody mistakes it for the device search's rate."""
    from posggym_baselines_b200.planning import (INTMCP, POTMMCP, MCTSConfig, OtherAgentMixturePolicy,
                                                 POTMMCPMetaPolicy)

    out = {}

    def episode(planner, key):
        rng = random.Random(7)
        obs, _ = env.reset(seed=300)
        planner.reset()
        sims = []
        for _ in range(decisions):
            a = planner.step(obs["0"])
            if planner.step_statistics["num_sims"] == planner.step_statistics["num_sims"]:
                sims.append(planner.step_statistics["num_sims"])
            obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
            if done:
                break
        planner.close()
        out[key] = sum(sims) / max(1, len(sims))

    try:
        cfg = MCTSConfig(discount=0.99, search_time_limit=0.1, c=1.25, truncated=True, action_selection="pucb",
                         state_belief_only=False, seed=0)
        types = {k: _TiltedPolicy(model, "1", k, f) for k, f in (("pusher", 1), ("idler", 0))}
        ego = {k: _TiltedPolicy(model, "0", k, f) for k, f in (("fast", 1), ("careful", 2))}
        meta = POTMMCPMetaPolicy.get_uniform_meta_policy({e: {o: 0.0 for o in types} for e in ego})
        episode(POTMMCP(model, "0", cfg, {"1": OtherAgentMixturePolicy(model, "1", types)},
                        POTMMCPMetaPolicy(model, "0", ego, meta)), "potmmcp_sims_per_decision")
        out["potmmcp_config"] = "Driving-v1, 2 other-agent types x 2 ego policies, value leaves, pucb, 0.1 s"
        icfg = MCTSConfig(discount=0.95, search_time_limit=0.2, c=1.414, truncated=False, action_selection="ucb",
                          state_belief_only=False, seed=0)
        episode(INTMCP.initialize(model, "0", icfg, nesting_level=2, search_policies=None),
                "intmcp_level2_sims_per_decision")
        out["intmcp_level2_config"] = "Driving-v1, nesting_level 2 (three trees deep), random search policies, ucb, 0.2 s"
    except Exception as ex:
        out["error"] = repr(ex)
    return out


def bench_models(dev, rank, world, seed, barrier, steps: int):
    """``model.step`` of the two other in-scope models as roofline-shaped sections: the same random-action
    depth-50 sweep (Philox initial states / joint actions, every output written) at 2^22 envs per GPU --
    per-launch working set 214 / 281 MB, larger than the 126 MB L2."""
    import torch
    import torch.distributed as dist

    from posggym_baselines_b200.envs.lbf import BatchedLBF
    from posggym_baselines_b200.envs.predator_prey import BatchedPredatorPrey

    peak, peak_src = load_peaks()
    counters = load_counters()
    n = 1 << 22
    out = {}
    for name, env, kernel in (("PredatorPrey-v0", BatchedPredatorPrey(device=dev), "posg::generic::step_fast_kernel<posg::PPSimT<true>, true, true>"),
                              ("LevelBasedForaging-v3", BatchedLBF(device=dev), "posg::generic::step_fast_kernel<posg::LBFSim, true, true>")):
        bytes_per = 8 + 8 + 16 * env.obs_words + 16 + 1 + 2   # state in; state, obs, f64 rewards, flags, actions out
        st0 = env.sample_initial_state(n, seed=seed, particle_offset=rank * n)
        live = torch.zeros(1, dtype=torch.int64, device=dev)
        bufs = [env.step(st0, seed=seed, step_idx=0, particle_offset=rank * n, live_count=live) for _ in range(2)]

        def sweep():
            cur = st0
            for k in range(DEPTH):
                o = bufs[k & 1]
                env.step(cur, seed=seed, step_idx=k, particle_offset=rank * n, out=o, live_count=live)
                cur = o.state

        for _ in range(3):
            sweep()
        barrier()
        live.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            sweep()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        lv = live.clone()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(lv, op=dist.ReduceOp.SUM)
        launch_ms = ms.item() / (steps * DEPTH)
        achieved = bytes_per * n / (launch_ms / 1e3) / 1e9
        c = counters.get(kernel, {})
        out[name] = {
            "env_steps_per_s": int(lv.item()) / (ms.item() / 1e3), "padded_env_steps_per_s": world * n * DEPTH * steps / (ms.item() / 1e3),
            "envs_per_gpu": n, "depth": DEPTH,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (c.get("dram_bytes_per_env_step") * n) if c.get("dram_bytes_per_env_step") else None,
                         "kernel": kernel, "bytes_per_env_step": bytes_per, "launch_ms": launch_ms, "peak_source": peak_src,
                         "thread_instructions_per_env_step": c.get("thread_instructions_per_env_step"),
                         "limiter": c.get("limiter")}}
        del bufs, st0
    return out


def bench_particle_sweep(env, dev, rank, seed, barrier):
    """BASELINE.json configs[1]: the Driving-v1 sweep from 1 K to 16 M particles (depth 50, one GPU's share).
    Below ~1 M particles the 50 launches are bound by the launch rate and the working set sits in L2."""
    import torch

    out = []
    for log2n in range(10, 25, 2):
        n = 1 << log2n
        st0 = env.sample_initial_state(n, seed=seed, particle_offset=rank * n)
        bufs = [env.step(st0, seed=seed, step_idx=0, particle_offset=rank * n) for _ in range(2)]
        reps = max(2, min(40, (1 << 22) // n))

        def sweep():
            cur = st0
            for k in range(DEPTH):
                o = bufs[k & 1]
                env.step(cur, seed=seed, step_idx=k, particle_offset=rank * n, out=o)
                cur = o.state

        sweep()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            sweep()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        row = {"particles": n, "ms_per_sweep": ms, "us_per_launch": 1e3 * ms / DEPTH,
               "padded_env_steps_per_s": n * DEPTH / (ms / 1e3),
               "algorithmic_gbs": BYTES_PER_ENV_STEP * n * DEPTH / (ms / 1e3) / 1e9,
               "working_set_mb": n * BYTES_PER_ENV_STEP / 1e6}
        if n <= (1 << 20):
            # small batches are bound by the host's launch rate (Python -> ctypes -> cudaLaunchKernel, ~12 us): the
            # same 50 launches captured once in a CUDA graph show what the device needs
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    sweep()
                g.replay()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(reps):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
                gms = e0.elapsed_time(e1) / reps
                row["cuda_graph"] = {"us_per_launch": 1e3 * gms / DEPTH, "padded_env_steps_per_s": n * DEPTH / (gms / 1e3)}
                del g
            except Exception as ex:   # pragma: no cover
                row["cuda_graph"] = {"error": repr(ex)}
        out.append(row)
        del bufs, st0
    return out


def bench_collector(dev, rank, world, seed, barrier):
    """BASELINE.json configs[4] with the policy in the loop: the rollout worker's step --
    policy.get_action_and_value on the current observations, envs.step with the sampled actions, store
    reward / done / value / log-prob (ppo/worker.py:135-224) -- for LevelBasedForaging-v3, 8 192 envs per GPU, 64
    steps: two launches per step (posg_policy_step on the tensor cores, posg_vec_step_embed) and the
    whole 64-step rollout captured in ONE CUDA graph.  Random-init network of the experiments' shape
    (trunk [64, 64], LSTM 64, heads [64]; exp_utils.py:101-104)."""
    import torch
    import torch.distributed as dist

    from posggym_baselines_b200.envs.lbf import LBFModel
    from posggym_baselines_b200.ppo.collector import DeviceCollector

    nn = torch.nn
    torch.manual_seed(0)
    sd = {}
    for name, mod in (("trunk.0", nn.Linear(300, 64)), ("trunk.2", nn.Linear(64, 64)), ("actor.0", nn.Linear(64, 64)),
                      ("actor.2", nn.Linear(64, 6)), ("critic.0", nn.Linear(64, 64)), ("critic.2", nn.Linear(64, 1))):
        sd[name + ".weight"], sd[name + ".bias"] = mod.weight.detach(), mod.bias.detach()
    lstm = nn.LSTM(64, 64)
    for k, v in lstm.state_dict().items():
        sd["lstm." + k] = v
    out = {"rollout_steps": 64, "launches_per_step": 2,
           "network": "PPOLSTMModel trunk [64, 64], LSTM 64, heads [64], random init; fp16 tensor-core operands, fp32 accumulate"}
    for label, nv in (("envs_8192_per_gpu", 8192), ("envs_131072_per_gpu", 131072)):
        col = DeviceCollector(LBFModel(device=dev), sd, nv, rollout_length=64, seed=seed, env_offset=rank * nv)
        col.capture()
        col.rollout()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            col.rollout()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[label] = {"us_per_step": 1e3 * ms.item() / 64, "env_steps_per_s": world * nv * 64 / (ms.item() / 1e3),
                      "agent_steps_per_s": 2 * world * nv * 64 / (ms.item() / 1e3)}
        del col
    return out


def bench_vec_env(dev, rank, world, seed, barrier, cpu: bool):
    """``envs.step`` of the PPO rollout worker (ppo/worker.py:176-216) for 64 rollout steps
    (exp_utils.py:81): one launch per step does model.step, TimeLimit, episode statistics,
    auto-reset and the one-hot float32 flattening (300 floats per agent for LBF-v3)."""
    import torch
    import torch.distributed as dist

    from posggym_baselines_b200.envs.driving import DrivingModel
    from posggym_baselines_b200.envs.lbf import LBFModel
    from posggym_baselines_b200.vector import BatchedVecEnv

    peak, _ = load_peaks()
    out = {"rollout_steps": 64}
    for name, model, nact in (("LevelBasedForaging-v3", LBFModel(device=dev), 6), ("Driving-v1", DrivingModel(device=dev), 5)):
        sec = {}
        for label, nv in (("envs_8192_per_gpu", 8192), ("envs_1048576_per_gpu", 1 << 20)):
            ve = BatchedVecEnv(model, nv, seed=seed, env_offset=rank * nv)
            ve.reset()
            acts = torch.randint(0, nact, (64, nv * 2), device=dev, dtype=torch.int64)
            for k in range(8):
                ve.step(acts[k])
            barrier()
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record()
            for k in range(64):          # the public call: envs.step(actions) -> (obs, rew, term, trunc, dones, infos)
                ve.step(acts[k])
            v1.record()
            barrier()
            vms = torch.tensor([v0.elapsed_time(v1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(vms, op=dist.ReduceOp.MAX)
            # state r+w, int64 actions, f32 one-hot obs, f32 rewards, term, trunc, done, episode r / l
            bytes_per = 8 + 8 + 16 + 2 * ve.flat_dim * 4 + 8 + 2 + 2 + 1 + 8 + 4
            gbs = bytes_per * nv * 64 / (vms.item() / 1e3) / 1e9
            sec[label] = {"env_steps_per_s": world * nv * 64 / (vms.item() / 1e3), "us_per_step": 1e3 * vms.item() / 64,
                          "hbm_gbs": gbs, "hbm_frac": gbs / peak, "bytes_per_env_step": bytes_per}
            sec["obs_dim"] = ve.flat_dim
            if nv == 8192:
                # the same 64 steps captured in ONE CUDA graph (envs.step_device; the step index lives
                # on the device): the eager call rate is bound by the ~16 us a Python -> ctypes launch
                # costs, the replay by the 8 us kernel.  A collector would capture its policy's kernels
                # between the steps in the same way.
                ve.enable_device_counter()
                for k in range(4):
                    ve.step_device(acts[k])
                barrier()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    for k in range(64):
                        ve.step_device(acts[k])
                graph.replay()
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(4):
                    graph.replay()
                g1.record()
                barrier()
                gms = torch.tensor([g0.elapsed_time(g1) / 4], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(gms, op=dist.ReduceOp.MAX)
                sec[label]["cuda_graph_rollout"] = {"us_per_step": 1e3 * gms.item() / 64,
                                                    "env_steps_per_s": world * nv * 64 / (gms.item() / 1e3)}
                del graph
            del ve
        if name == "LevelBasedForaging-v3":
            # SURVEY 8(f) N1, first half: envs.step + the policy's first trunk layer (Linear(300, 64) + Tanh,
            # ppo/network.py:171-178).  A: one-hot rows + torch (cuBLAS sgemm + tanh); B: no one-hot rows,
            # posg_vec_embed sums one weight column per observation field.
            sec["step_plus_trunk0"] = {}
            w = torch.randn(64, 300, device=dev) * 0.1
            bia = torch.zeros(64, device=dev)
            for label, nv in (("envs_8192_per_gpu", 8192), ("envs_1048576_per_gpu", 1 << 20)):
                res = {}
                for mode in ("one_hot_plus_torch_linear", "packed_embed"):
                    ve = BatchedVecEnv(model, nv, seed=seed, env_offset=rank * nv, write_obs=(mode != "packed_embed"))
                    ve.reset()
                    acts = torch.randint(0, nact, (32, nv * 2), device=dev, dtype=torch.int64)
                    hid = torch.empty((nv * 2, 64), dtype=torch.float32, device=dev)

                    def one(k):
                        o = ve.step(acts[k])[0]
                        if mode == "packed_embed":
                            ve.embed(w, bia, out=hid)
                        else:
                            torch.tanh(torch.nn.functional.linear(o, w, bia), out=hid)

                    for k in range(6):
                        one(k)
                    barrier()
                    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    v0.record()
                    for k in range(32):
                        one(k)
                    v1.record()
                    barrier()
                    vms = torch.tensor([v0.elapsed_time(v1)], dtype=torch.float64, device=dev)
                    if world > 1:
                        dist.all_reduce(vms, op=dist.ReduceOp.MAX)
                    res[mode] = {"us_per_step": 1e3 * vms.item() / 32, "env_steps_per_s": world * nv * 32 / (vms.item() / 1e3)}
                    del ve
                res["speedup"] = res["one_hot_plus_torch_linear"]["us_per_step"] / res["packed_embed"]["us_per_step"]
                sec["step_plus_trunk0"][label] = res
        out[name] = sec
    if cpu:
        # reference collector scale (BR checkpoints' config): 32 envs per worker stepped serially
        # (ppo/worker.py:51); one worker on one core here.
        from oracle import shims

        shims.install()
        import numpy as np

        from oracle import vec_env as V
        from oracle.lbf import LevelBasedForagingModel
        from posggym_baselines_b200.envs.env import Env

        def thunk():
            m = LevelBasedForagingModel()
            m.spec = type("S", (), {"max_episode_steps": 50, "id": "LevelBasedForaging-v3"})()
            return V.FlattenObservations(Env(m, 50))

        envs = V.StackEnv(V.RecordEpisodeStatistics(V.SyncVectorEnv([thunk] * 32)))
        envs.reset(seed=0)
        rng = np.random.default_rng(0)
        t0, k = time.perf_counter(), 0
        while time.perf_counter() - t0 < 3.0:
            envs.step(rng.integers(0, 6, size=64))
            k += 1
        out["cpu_reference"] = {"env_steps_per_s": 32 * k / (time.perf_counter() - t0), "cores": 1, "num_envs": 32,
                                "what": "reference vec-env stack restated on the CPU oracle (oracle/vec_env.py), "
                                        "LevelBasedForaging-v3, serial stepping as ppo/worker.py:51"}
    return out


_JSON_FD = None


def quiet_stdout():
    """Library chatter (e.g. "NCCL version ..." with NCCL_DEBUG=VERSION) must not share stdout with
    the ONE JSON line: fd 1 is pointed at stderr for the whole run, the line goes to the saved fd."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=1 << 24, help="particles per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-bind", action="store_true", help="multi-GPU: do not bind ranks to their GPU's NUMA node")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only-sweep", action="store_true", help="skip the rollout / vec-env / POMCP side sections")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from posggym_baselines_b200.envs.driving import BatchedDriving, StepOut

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # multi-GPU: every rank on the cores / memory of its GPU's NUMA node before any pinned buffer exists (at N = 1
    # the process stays unbound: the CPU baseline wants every core)
    host_binding = None
    if world > 1 and not args.no_bind:
        from posggym_baselines_b200.hostbind import bind_to_gpu_node

        host_binding = bind_to_gpu_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n, seed = args.particles, 2026
    env = BatchedDriving(device=dev)
    offset = rank * n  # Philox counters are global particle indices: results independent of N
    st0 = env.sample_initial_state(n, seed=seed, particle_offset=offset)
    mk = lambda: StepOut(  # noqa: E731
        torch.empty((n, 2), dtype=torch.int32, device=dev), torch.empty((n, 2), dtype=torch.int64, device=dev),
        torch.empty((n, 2), dtype=torch.float64, device=dev), torch.empty((n,), dtype=torch.uint8, device=dev),
        torch.empty((n, 2), dtype=torch.uint8, device=dev))
    bufs = [mk(), mk()]
    live = torch.zeros(1, dtype=torch.int64, device=dev)

    def sweep():
        cur = st0
        for k in range(DEPTH):
            out = bufs[k & 1]
            env.step(cur, seed=seed, step_idx=k, particle_offset=offset, out=out, live_count=live)
            cur = out.state

    for _ in range(max(3, args.warmup)):
        sweep()
    barrier()
    live.zero_()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        sweep()
    e1.record()
    barrier()
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    live_total = live.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(live_total, op=dist.ReduceOp.SUM)
    ms = float(ms.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    executed = int(live_total.item())
    padded = n * DEPTH * args.steps * world
    value = executed / (ms / 1e3)
    launches = args.steps * DEPTH
    launch_ms = ms / launches
    peak, peak_src = load_peaks()
    achieved = BYTES_PER_ENV_STEP * n / (launch_ms / 1e3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu capture
    # (profiles/step_kernel_traffic.json), scaled to this batch size
    traffic = None
    tp = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_per_launch"] * (n / tj["particles"])
        except Exception:
            traffic = None

    # ---- fused rollout kernel (not HBM-bound; reported beside the step sweep) ----
    rl = None
    r_live = torch.zeros(1, dtype=torch.int64, device=dev)
    for _ in range(0 if args.only_sweep else 3):
        env.rollout(st0, 0, DEPTH, 0.95, seed=seed, particle_offset=offset)
    barrier()
    r_live.zero_()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(1 if args.only_sweep else args.steps):
        env.rollout(st0, 0, DEPTH, 0.95, seed=seed, particle_offset=offset, live_count=r_live)
    r1.record()
    barrier()
    rms = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
    rl_live = r_live.clone()
    if world > 1:
        dist.all_reduce(rms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rl_live, op=dist.ReduceOp.SUM)
    rl_steps = 1 if args.only_sweep else args.steps
    rl = {"rollouts_per_s": n * world * rl_steps / (rms.item() / 1e3),
          "env_steps_per_s": int(rl_live.item()) / (rms.item() / 1e3),
          "ms_per_batch": rms.item() / rl_steps, "depth": DEPTH, "ego_agent": 0,
          "note": "fused MCTS._rollout kernel: stops at ego termination; issue-bound, ~16 B of HBM per rollout"}
    # this kernel moves ~0.4 B of HBM per env-step; its bound is the issue rate (148 SMs x 4
    # schedulers x SM clock warp-instructions/s).  Warp-instructions per env-step come from the committed ncu
    # capture of this kernel (profiles/kernel_counters.json); the rate is measured here.
    rc = load_counters().get("posg::driving::rollout_kernel", {})
    if rc.get("warp_instructions_per_env_step"):
        sm_hz = (rc.get("sm_mhz") or 1965.0) * 1e6
        issue_peak = 148 * 4 * sm_hz
        issue_ach = rc["warp_instructions_per_env_step"] * rl["env_steps_per_s"] / max(1, world)
        rl["roofline"] = {"bound": "issue", "achieved": issue_ach, "peak": issue_peak, "unit": "warp-instructions/s",
                          "frac": issue_ach / issue_peak, "kernel": "posg::driving::rollout_kernel",
                          "warp_instructions_per_env_step": rc["warp_instructions_per_env_step"],
                          "lanes_active": rc.get("avg_active_threads_per_warp"),
                          "traffic": rc.get("dram_bytes_per_launch")}

    # ---- e2e: C-ABI with pinned HOST buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty((n, 2), dtype=torch.int32).pin_memory()
        h_in.copy_(st0.cpu())
        outs = (torch.empty((n, 2), dtype=torch.int32).pin_memory().numpy(),
                torch.empty((n, 2), dtype=torch.int64).pin_memory().numpy(),
                torch.empty((n, 2), dtype=torch.float64).pin_memory().numpy(),
                torch.empty((n,), dtype=torch.uint8).pin_memory().numpy())
        h_np = h_in.numpy()
        e_steps = max(2, min(args.steps, 5))
        for _ in range(2):
            env.sweep_host(h_np, DEPTH, seed=seed, particle_offset=offset, outs=outs)
        barrier()
        w0 = time.perf_counter()
        e_live = 0
        for _ in range(e_steps):
            _, lv = env.sweep_host(h_np, DEPTH, seed=seed, particle_offset=offset, outs=outs)
            e_live += lv
        torch.cuda.synchronize()
        w = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
        lv_t = torch.tensor([e_live], dtype=torch.int64, device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
            dist.all_reduce(lv_t, op=dist.ReduceOp.SUM)
        e2e = {"value": int(lv_t.item()) / w.item(), "unit": UNIT,
               "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * (8 + 16 + 16 + 1) + 8,
               "ms_per_step": 1e3 * w.item() / e_steps, "steps": e_steps, "host_binding": host_binding,
               "api": "posg_driving_sweep_host (pinned host state in; last step's state/obs/rewards/flags out)"}
        # compact result mode: rewards as one code byte per env (decoded exactly on the host from the reward table
        # and the returned state: posg_driving_decode_rewards_host): 26 instead of 41 B/env device -> host
        try:
            c_outs = (outs[0], outs[1], torch.empty((n,), dtype=torch.uint8).pin_memory().numpy(), outs[3])
            for _ in range(2):
                env.sweep_host_compact(h_np, DEPTH, seed=seed, particle_offset=offset, outs=c_outs)
            barrier()
            w0 = time.perf_counter()
            c_live = 0
            for _ in range(e_steps):
                _, lv = env.sweep_host_compact(h_np, DEPTH, seed=seed, particle_offset=offset, outs=c_outs)
                c_live += lv
            torch.cuda.synchronize()
            wc = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
            lc_t = torch.tensor([c_live], dtype=torch.int64, device=dev)
            if world > 1:
                dist.barrier()
                dist.all_reduce(wc, op=dist.ReduceOp.MAX)
                dist.all_reduce(lc_t, op=dist.ReduceOp.SUM)
            e2e["compact_results"] = {"value": int(lc_t.item()) / wc.item(), "unit": UNIT, "h2d_bytes_per_step": n * 8,
                                      "d2h_bytes_per_step": n * (8 + 16 + 1 + 1) + 8, "ms_per_step": 1e3 * wc.item() / e_steps,
                                      "api": "posg_driving_sweep_host_compact (rewards as reward-table codes)"}
        except Exception as ex:
            e2e["compact_results"] = {"error": repr(ex)}

    # ---- CPU baseline (rank 0, N=1 only): oracle port on the host cores ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import cpu_bench

        cores = os.cpu_count() or 1
        r = cpu_bench.driving_sweep(budget_s=args.cpu_seconds, depth=DEPTH, procs=cores, seed=0)
        try:
            rc = cpu_bench.driving_sweep_compiled(budget_s=min(5.0, args.cpu_seconds), depth=DEPTH, threads=cores, seed=0)
            compiled = {"value": rc["env_steps_per_s"], "unit": UNIT, "cores": cores,
                        "what": "C restatement of the model (oracle/c/driving_oracle.c, gcc -O2), one thread per core"}
        except Exception as ex:
            compiled = {"error": repr(ex)}
        cpu = {"value": r["env_steps_per_s"], "unit": UNIT, "cores": cores, "kind": "port", "compiled_port": compiled,
               "sample": f"{r['env_steps']} env-steps of the same random-action depth-{DEPTH} Driving-v1 sweep in "
                         f"{r['seconds']:.1f} s, one Python process per core (the reference's model is pure Python)"}

    # ---- PPO vector-env collection (BASELINE.json configs[4]: LevelBasedForaging-v3, 65 536 envs over
    # 8 GPUs = 8 192 per GPU, 64 rollout steps, uniform actions; plus a large batch that shows the
    # HBM-bound regime, and Driving-v1 for comparison).
    vec = None
    try:
        if args.only_sweep:
            raise RuntimeError("skipped (--only-sweep)")
        vec = bench_vec_env(dev, rank, world, seed, barrier, cpu=(rank == 0 and world == 1 and not args.no_cpu))
    except Exception as ex:
        vec = {"error": repr(ex)}

    collector = None
    try:
        if args.only_sweep:
            raise RuntimeError("skipped (--only-sweep)")
        collector = bench_collector(dev, rank, world, seed, barrier)
    except Exception as ex:
        collector = {"error": repr(ex)}

    # ---- the other two models' step kernels and the particle sweep of configs[1] ----
    models, psweep = None, None
    try:
        if args.only_sweep:
            raise RuntimeError("skipped (--only-sweep)")
        models = bench_models(dev, rank, world, seed, barrier, max(2, min(args.steps, 4)))
        if world == 1:
            psweep = bench_particle_sweep(env, dev, rank, seed, barrier)
    except Exception as ex:
        models = {"error": repr(ex)}

    # ---- POMCP simulations per 0.1 s decision (second half of BASELINE.json's metric) ----
    pomcp = None
    try:
        if args.only_sweep:
            raise RuntimeError("skipped (--only-sweep)")
        pomcp = bench_pomcp(dev, rank, world, cpu=(rank == 0 and world == 1 and not args.no_cpu))
    except Exception as ex:  # the headline number must not be lost to a planner problem
        pomcp = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(n, seed), "padded_env_steps_per_s": padded / (ms / 1e3),
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "posg::driving::step_fast_kernel<true, true>", "bytes_per_env_step": BYTES_PER_ENV_STEP,
                         "launch_ms": launch_ms, "peak_source": peak_src},
            "cpu_baseline": cpu, "clocks": clocks, "rollout_fused": rl, "models": models, "particle_sweep": psweep,
            "pomcp": pomcp, "vec_env": vec, "collector": collector,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
