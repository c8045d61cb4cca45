"""Device pieces of the PPO learner's batch post-processing against the reference's tensor code."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference_gae(rewards_buf, dones_buf, values_buf, gamma, gae_lambda):
    """The loop of PPOTrainer.update (ppo/core.py:226-244), verbatim semantics."""
    advantages = torch.zeros_like(rewards_buf)
    lastgaelam = 0
    for t in reversed(range(rewards_buf.shape[0])):
        nextnonterminal = 1.0 - dones_buf[t + 1]
        nextvalues = values_buf[t + 1]
        delta = rewards_buf[t] + gamma * nextvalues * nextnonterminal - values_buf[t]
        advantages[t] = lastgaelam = delta + gamma * gae_lambda * nextnonterminal * lastgaelam
    return advantages, advantages + values_buf[:-1]


@pytest.mark.parametrize("shape", [(64, 8192, 2), (16, 33), (1, 5, 2), (128, 1000)])
def test_gae_bit_identical_to_reference_loop(shape):
    from posggym_baselines_b200.ppo import compute_gae

    g = torch.Generator().manual_seed(sum(shape))
    t, batch = shape[0], shape[1:]
    rewards = (torch.rand((t, *batch), generator=g) * (torch.rand((t, *batch), generator=g) < 0.1)).cuda()
    dones = (torch.rand((t + 1, *batch), generator=g) < 0.05).float().cuda()
    values = torch.randn((t + 1, *batch), generator=g).cuda()
    for gamma, lam in ((0.99, 0.95), (0.95, 1.0), (1.0, 0.9)):
        adv_ref, ret_ref = reference_gae(rewards, dones, values, gamma, lam)
        adv, ret = compute_gae(rewards, dones, values, gamma, lam)
        assert adv.shape == rewards.shape
        assert torch.equal(adv.view(torch.int32), adv_ref.view(torch.int32)), (adv - adv_ref).abs().max()
        assert torch.equal(ret.view(torch.int32), ret_ref.view(torch.int32))


def test_gae_empty():
    from posggym_baselines_b200.ppo import compute_gae

    adv, ret = compute_gae(torch.zeros((0, 4), device="cuda"), torch.zeros((1, 4), device="cuda"),
                           torch.zeros((1, 4), device="cuda"), 0.99, 0.95)
    assert adv.shape == (0, 4) and ret.shape == (0, 4)
