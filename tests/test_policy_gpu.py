"""The fused tensor-core policy step (posg_policy_step, csrc/policy_kernels.cu) against a plain fp32
PyTorch restatement of the same op -- PPOLSTMModel.get_action_and_value of the reference
(ppo/network.py:212-268,305-330) with trunk_sizes [64, 64], lstm_size 64, head_sizes [64]
(baseline_exps/exp_utils.py:101-104).  Operands are fp16 on the tensor cores (fp32 accumulation and
epilogues): tolerance 5e-3 absolute on logits / value / log-prob / entropy / h / c, stated per assert."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class RefNet(torch.nn.Module):
    """The reference module's layers under the reference's parameter names."""

    def __init__(self, obs_dim, n_actions):
        super().__init__()
        nn = torch.nn
        self.trunk = nn.Sequential(nn.Linear(obs_dim, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh())
        self.lstm = nn.LSTM(64, 64, num_layers=1, batch_first=False)
        self.actor = nn.Sequential(nn.Linear(64, 64), nn.Tanh(), nn.Linear(64, n_actions))
        self.critic = nn.Sequential(nn.Linear(64, 64), nn.Tanh(), nn.Linear(64, 1))

    def get_action_and_value(self, x, lstm_state, done, action):
        hidden = self.trunk(x)
        keep = (1.0 - done).view(1, -1, 1)
        out, state = self.lstm(hidden.unsqueeze(0), (keep * lstm_state[0], keep * lstm_state[1]))
        out = out.squeeze(0)
        logits = self.actor(out)
        probs = torch.distributions.Categorical(logits=logits)
        return logits, probs.log_prob(action), probs.entropy(), self.critic(out).squeeze(-1), state


@pytest.mark.parametrize("n,n_actions", [(128, 6), (1000, 5), (16384, 6)])
def test_policy_step_matches_fp32_torch(n, n_actions):
    from posggym_baselines_b200.ppo import FusedPolicy

    torch.manual_seed(n)
    dev = torch.device("cuda")
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        net = RefNet(300, n_actions).to(dev)
        with torch.no_grad():                       # trained-looking magnitudes, biases included
            for p in net.parameters():
                p.mul_(1.5).add_(0.05 * torch.randn_like(p))
        pol = FusedPolicy(net.state_dict(), device=dev)
        x = (torch.rand(n, 300, device=dev) < 0.1).float()
        h = torch.randn(1, n, 64, device=dev) * 0.5
        c = torch.randn(1, n, 64, device=dev)
        done = (torch.rand(n, device=dev) < 0.2).float()
        action = torch.randint(0, n_actions, (n,), device=dev)
        with torch.no_grad():
            logits, logp, ent, val, (h1, c1) = net.get_action_and_value(x, (h, c), done, action)
            x0 = net.trunk[1](net.trunk[0](x)).contiguous()
        hh, cc = h[0].clone(), c[0].clone()
        out = pol.step(x0, (hh, cc), done, action=action, want_logits=True)
        torch.cuda.synchronize()
        tol = 5e-3
        assert (out["logits"] - logits).abs().max().item() < tol
        assert (out["value"] - val).abs().max().item() < tol
        assert (out["logprob"] - logp).abs().max().item() < tol
        assert (out["entropy"] - ent).abs().max().item() < tol
        assert (hh - h1[0]).abs().max().item() < tol and (cc - c1[0]).abs().max().item() < tol
        assert torch.equal(out["action"], action)
        # sampling: Philox-driven inverse CDF -- reproducible, and distributed like Categorical(logits)
        hh2, cc2 = h[0].clone(), c[0].clone()
        s1 = pol.step(x0, (hh2, cc2), done, seed=5, step_idx=3)["action"]
        hh3, cc3 = h[0].clone(), c[0].clone()
        s2 = pol.step(x0, (hh3, cc3), done, seed=5, step_idx=3)["action"]
        assert torch.equal(s1, s2) and int(s1.min()) >= 0 and int(s1.max()) < n_actions
        if n >= 16384:
            counts = torch.zeros(n_actions, device=dev)
            expect = torch.softmax(logits, -1).sum(0)
            for k in range(8):
                a = pol.step(x0, (h[0].clone(), c[0].clone()), done, seed=9, step_idx=k)["action"]
                counts += torch.bincount(a, minlength=n_actions).float()
            freq, p = (counts / (8 * n)).cpu().numpy(), (expect / n).cpu().numpy()
            assert np.abs(freq - p).max() < 4 * np.sqrt(p.max() / (8 * n)) + 2e-3, (freq, p)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_sass_uses_the_tensor_core_path():
    """UTCHMMA (tcgen05.mma kind::f16) and LDTM (tcgen05.ld) are in the shipped library."""
    import subprocess

    from posggym_baselines_b200 import _lib

    sass = subprocess.run(["cuobjdump", "-sass", _lib.library_path()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTCBAR" in sass


def test_device_collector_matches_stepwise_fp32_reference():
    """DeviceCollector (embed + tensor-core policy + env step per rollout step) against the same loop written with
    the public vec env (one-hot rows) and the fp32 PyTorch module, teacher-forced with the collector's sampled
    actions: rewards / dones / env states identical, values / log-probs / entropies within 5e-3."""
    from posggym_baselines_b200.envs.lbf import LBFModel
    from posggym_baselines_b200.ppo.collector import DeviceCollector
    from posggym_baselines_b200.vector import BatchedVecEnv

    torch.manual_seed(3)
    dev = torch.device("cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = LBFModel(device=dev)
    net = RefNet(300, 6).to(dev)
    n_envs, T = 256, 24
    col = DeviceCollector(model, net.state_dict(), n_envs, rollout_length=T, seed=17)
    roll = col.rollout()
    torch.cuda.synchronize()
    ref_env = BatchedVecEnv(model, n_envs, seed=17)
    obs, _ = ref_env.reset()
    R = n_envs * 2
    h = torch.zeros(1, R, 64, device=dev)
    c = torch.zeros(1, R, 64, device=dev)
    done = torch.zeros(R, device=dev)
    for t in range(T):
        assert torch.equal(roll["dones"][t].float(), done)
        assert torch.equal(roll["states"][t], ref_env._state)
        with torch.no_grad():
            logits, logp, ent, val, (h, c) = net.get_action_and_value(obs, (h, c), done, roll["actions"][t])
        assert (roll["logprobs"][t] - logp).abs().max().item() < 5e-3
        assert (roll["values"][t] - val).abs().max().item() < 5e-3
        assert (roll["entropies"][t] - ent).abs().max().item() < 5e-3
        obs, rew, term, trunc, dones, _ = ref_env.step(roll["actions"][t])
        assert torch.equal(roll["rewards"][t], rew)
        done = (term | trunc).float()
    assert (col.h - h[0]).abs().max().item() < 2e-2          # 24 recurrent steps of fp16-operand rounding


def test_device_collector_graph_replays_are_fresh_rollouts():
    from posggym_baselines_b200.envs.lbf import LBFModel
    from posggym_baselines_b200.ppo.collector import DeviceCollector

    dev = torch.device("cuda")
    net = RefNet(300, 6).to(dev)
    a = DeviceCollector(LBFModel(device=dev), net.state_dict(), 512, rollout_length=16, seed=5)
    b = DeviceCollector(LBFModel(device=dev), net.state_dict(), 512, rollout_length=16, seed=5)
    a.capture()                      # = one eager warm-up rollout + the capture
    b.rollout_eager()                # the same first rollout, eagerly
    r1 = {k: v.clone() for k, v in a.rollout().items()}      # second rollout: graph replay
    b.rollout_eager()                                         # second rollout: eager
    torch.cuda.synchronize()
    for k in ("actions", "rewards", "dones", "states"):
        assert torch.equal(r1[k], getattr(b, k)[:r1[k].shape[0]]), k   # (states keeps one more row: the next rollout's first)
    r2 = a.rollout()
    torch.cuda.synchronize()
    assert not torch.equal(r1["actions"], r2["actions"])      # a replay continues the envs with new draws


@pytest.mark.parametrize("n", [1, 33, 1000])
def test_policy_eval_host_matches_fp32_torch(n):
    """posg_policy_eval_host -- the planners' one-call network evaluation with host rows (trunk layer 0 as an
    index-row sum, then the tensor-core kernel): new LSTM state, logits and value within 5e-3 of the fp32 module fed
    the one-hot rows those indices stand for; out-of-range indices contribute nothing."""
    from posggym_baselines_b200.ppo import FusedPolicy

    torch.manual_seed(100 + n)
    dev = torch.device("cuda")
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        net = RefNet(124, 5).to(dev)
        with torch.no_grad():
            for p in net.parameters():
                p.mul_(1.5).add_(0.05 * torch.randn_like(p))
        pol = FusedPolicy(net.state_dict(), device=dev)
        rng = np.random.default_rng(n)
        k = 22
        idx = np.stack([rng.choice(124, size=k, replace=False) for _ in range(n)]).astype(np.int32)
        x = torch.zeros(n, 124, device=dev)
        x[torch.arange(n, device=dev).view(-1, 1), torch.from_numpy(idx.astype(np.int64)).to(dev)] = 1.0
        idx_padded = np.concatenate([idx, np.full((n, 2), -1, np.int32)], 1)
        idx_padded[:, -1] = 124                                  # both pads are out of range
        h = (torch.randn(n, 64) * 0.5).numpy()
        c = torch.randn(n, 64).numpy()
        with torch.no_grad():
            logits, _, _, val, (h1, c1) = net.get_action_and_value(
                x, (torch.from_numpy(h).to(dev).unsqueeze(0), torch.from_numpy(c).to(dev).unsqueeze(0)),
                torch.zeros(n, device=dev), torch.zeros(n, dtype=torch.int64, device=dev))
        h2, c2, lg, pr, v = pol.eval_host(idx_padded, h, c)
        tol = 5e-3
        assert np.abs(pr - torch.softmax(logits, -1).cpu().numpy()).max() < tol and np.abs(pr.sum(1) - 1).max() < 1e-5
        assert np.abs(lg - logits.cpu().numpy()).max() < tol and np.abs(v - val.cpu().numpy()).max() < tol
        assert np.abs(h2 - h1[0].cpu().numpy()).max() < tol and np.abs(c2 - c1[0].cpu().numpy()).max() < tol
        h3, c3, lg3, _, v3 = pol.eval_host(idx_padded, h, c)        # the staging is reused: same answer
        assert np.array_equal(lg, lg3) and np.array_equal(h2, h3) and np.array_equal(v, v3)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
