"""Sequence chunking of rollout batches (posggym_baselines_b200.ppo.seq: posg_seq_chunks / posg_seq_gather,
csrc/seq_kernels.cu) against the per-env loops of the reference (restated in oracle/ppo_utils.py and, when
/root/reference is importable, the reference module itself), and the per-policy batch views
(posggym_baselines_b200.ppo.batch) against the reference's functions."""
import pytest
import torch

from oracle import ppo_utils as O
from posggym_baselines_b200.ppo.seq import get_seq_chunks, split_and_pad_batch, split_lstm_state_batch


def cases():
    g = torch.Generator().manual_seed(0)
    out = []
    for T, B, p, L in ((64, 37, 0.05, 16), (16, 5, 0.3, 4), (8, 3, 0.0, 8), (50, 11, 0.1, 7), (5, 4, 1.0, 3), (12, 6, 0.2, 20)):
        dones = (torch.rand((T, B), generator=g) < p).float()
        out.append((dones, L))
    d = torch.zeros((10, 2)); d[0, 0] = 1; d[9, 1] = 1      # done at the first / last step
    out.append((d, 4))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(7))
def test_chunks_and_padding_match_reference_loops(case):
    dones, L = cases()[case]
    T, B = dones.shape
    ref_idx, ref_over = O.get_seq_idxs(dones, L)
    ch = get_seq_chunks(dones.cuda(), L)
    got_idx, got_over = SeqChunksCpu(ch).to_lists(B)
    assert [[c.tolist() for c in e] for e in got_idx] == [[c.tolist() for c in e] for e in ref_idx]
    assert [[c.tolist() for c in e] for e in got_over] == [[c.tolist() for c in e] for e in ref_over]
    g = torch.Generator().manual_seed(1)
    batch = torch.randn((T, B, 2, 3), generator=g)
    for pad in (0.0, -1.5):
        assert torch.equal(split_and_pad_batch(batch.cuda(), ch, L, pad).cpu(), O.split_and_pad_batch(batch, ref_idx, L, pad))
    batch1 = torch.randn((T + 1, B, 2), generator=g)
    assert torch.equal(split_and_pad_batch(batch1.cuda(), ch, L, 0.0, overlapping=True).cpu(),
                       O.split_and_pad_batch(batch1, ref_over, L + 1, 0.0))
    ints = torch.randint(0, 6, (T, B, 2), generator=g)
    assert torch.equal(split_and_pad_batch(ints.cuda(), ch, L, 0.0).cpu(), O.split_and_pad_batch(ints, ref_idx, L, 0.0))
    lstm = torch.randn((T, 1, B, 2, 4), generator=g)
    assert torch.equal(split_lstm_state_batch(lstm.cuda(), ch).cpu(), O.split_lstm_state_batch(lstm, ref_idx))
    # every entry of the batch in one launch (mixed dtypes / row sizes / overlapping)
    from posggym_baselines_b200.ppo.seq import split_and_pad_batches

    bytes_ = torch.randint(0, 2, (T, B), generator=g).to(torch.uint8)
    outs = split_and_pad_batches([batch.cuda(), batch1.cuda(), ints.cuda(), bytes_.cuda()], ch, L, [0.0, -1.5, 0.0, 0.0],
                                 [False, True, False, False])
    assert torch.equal(outs[0].cpu(), O.split_and_pad_batch(batch, ref_idx, L, 0.0))
    assert torch.equal(outs[1].cpu(), O.split_and_pad_batch(batch1, ref_over, L + 1, -1.5))
    assert torch.equal(outs[2].cpu(), O.split_and_pad_batch(ints, ref_idx, L, 0.0))
    assert torch.equal(outs[3].cpu(), O.split_and_pad_batch(bytes_, ref_idx, L, 0.0))


def test_oracle_restatement_matches_reference_module(oracle_env):
    """Pins oracle/ppo_utils.py to the reference's own functions (build container only)."""
    if not oracle_env["have_reference"]:
        pytest.skip("/root/reference not present")
    try:
        from posggym_baselines.ppo import utils as R
    except Exception as ex:      # the module imports more of posggym than the shims provide
        pytest.skip(f"reference ppo.utils not importable: {ex!r}")
    for dones, L in cases():
        a, b = R.get_seq_idxs(dones, L), O.get_seq_idxs(dones, L)
        assert [[c.tolist() for c in e] for e in a[0]] == [[c.tolist() for c in e] for e in b[0]]
        assert [[c.tolist() for c in e] for e in a[1]] == [[c.tolist() for c in e] for e in b[1]]
        batch = torch.randn((dones.shape[0], dones.shape[1], 2, 3))
        assert torch.equal(R.split_and_pad_batch(batch, a[0], L, -2.0), O.split_and_pad_batch(batch, b[0], L, -2.0))
        lstm = torch.randn((dones.shape[0], 1, dones.shape[1], 2, 4))
        assert torch.equal(R.split_lstm_state_batch(lstm, a[0]), O.split_lstm_state_batch(lstm, b[0]))


@pytest.mark.gpu
def test_chunks_on_device_at_scale():
    """65 536 envs x 64 steps on the GPU (the reference's loop takes seconds here): same chunks as
    the CPU run of the same code, spot-checked against the loop oracle on a slice."""
    g = torch.Generator().manual_seed(3)
    dones = (torch.rand((64, 65536), generator=g) < 0.03).float()
    ch_gpu = get_seq_chunks(dones.cuda(), 16)
    # chunk count from first principles: piece starts (step 0, done steps) + the cuts every 16 steps inside pieces
    assert int(ch_gpu.length.sum()) == 64 * 65536 and int(ch_gpu.length.max()) <= 16 and ch_gpu.longest == 16
    assert bool((ch_gpu.env[1:] >= ch_gpu.env[:-1]).all())
    ref_idx, _ = O.get_seq_idxs(dones[:, :200], 16)
    sub = get_seq_chunks(dones[:, :200].cuda(), 16)
    got, _ = SeqChunksCpu(sub).to_lists(200)
    assert [[c.tolist() for c in e] for e in got] == [[c.tolist() for c in e] for e in ref_idx]
    batch = torch.randn((64, 65536, 2), generator=g).cuda()
    out = split_and_pad_batch(batch, ch_gpu, 16)
    assert out.shape == (16, ch_gpu.env.numel(), 2)


def SeqChunksCpu(ch):
    from posggym_baselines_b200.ppo.seq import SeqChunks

    return SeqChunks(ch.env.cpu(), ch.start.cpu(), ch.length.cpu())


class _Cfg:
    """The accessors of the reference's PPOConfig that the batch utilities read."""

    num_agents = 2
    ids = ["P0", "P1", "P2"]

    def get_all_policy_ids(self):
        return list(self.ids)

    def get_policy_idx(self, pid):
        return self.ids.index(pid)

    def get_policy_partner_distribution(self, pid):
        return {"P0": {"P0": 0.5, "P1": 0.5}, "P1": {"P0": 1.0}, "P2": {"P0": 0.3, "P1": 0.3, "P2": 0.4}}[pid]


def _fake_batch(L=5, B=12, N=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    pol = torch.randint(0, 3, (1, B, N), generator=g).expand(L, B, N).contiguous()
    return {"obs": torch.randn((L, B, N, 7), generator=g), "actions": torch.randint(0, 5, (L, B, N), generator=g),
            "policy_idxs": pol, "initial_lstm_states": (torch.randn((1, B, N, 4), generator=g), torch.randn((1, B, N, 4), generator=g)),
            "policy_stats": {"P0": {"mean_r": 1.0, "max_l": 3.0}, "P1": {"mean_r": 2.0}}}


def test_per_policy_batch_views():
    from posggym_baselines_b200.ppo import combine_batches, filter_policy_batches_by_partner_dist, split_batch_by_policy

    cfg, batch = _Cfg(), _fake_batch()
    split = split_batch_by_policy(batch, cfg)
    pol = batch["policy_idxs"][0]
    total = 0
    for pid in cfg.ids:
        mask = pol == cfg.get_policy_idx(pid)
        pb = split[pid]
        assert pb["obs"].shape == (5, int(mask.sum()), 7) and torch.equal(pb["obs"], batch["obs"][:, mask])
        assert pb["initial_lstm_states"][0].shape == (1, int(mask.sum()), 4)
        # partner of agent i = the other agent's policy
        assert torch.equal(pb["partner_policy_idxs"][0, :, 0], pol.flip(1)[mask])
        total += int(mask.sum())
    assert total == pol.numel()
    filt = filter_policy_batches_by_partner_dist(split, cfg)
    p1 = filt["P1"]          # P1 only trains against P0
    assert bool((p1["partner_policy_idxs"][0] == 0).all()) and p1["obs"].shape[1] <= split["P1"]["obs"].shape[1]
    comb = combine_batches([_fake_batch(seed=1), _fake_batch(seed=2)], cfg)
    assert comb["obs"].shape == (5, 24, 2, 7) and comb["initial_lstm_states"][1].shape == (1, 24, 2, 4)
    assert float(comb["policy_stats"]["P0"]["mean_r"]) == 1.0 and float(comb["policy_stats"]["P0"]["max_l"]) == 3.0


def test_per_policy_batch_views_match_reference_module(oracle_env):
    if not oracle_env["have_reference"]:
        pytest.skip("/root/reference not present")
    try:
        from posggym_baselines.ppo import utils as R
    except Exception as ex:
        pytest.skip(f"reference ppo.utils not importable: {ex!r}")
    from posggym_baselines_b200.ppo import combine_batches, filter_policy_batches_by_partner_dist, split_batch_by_policy

    cfg = _Cfg()
    for seed in range(3):
        batch = _fake_batch(seed=seed)
        a, b = R.split_batch_by_policy(batch, cfg), split_batch_by_policy(batch, cfg)
        fa, fb = R.filter_policy_batches_by_partner_dist(a, cfg), filter_policy_batches_by_partner_dist(b, cfg)
        for x, y in ((a, b), (fa, fb)):
            assert set(x) == set(y)
            for pid in x:
                for k in x[pid]:
                    if k == "initial_lstm_states":
                        assert torch.equal(x[pid][k][0], y[pid][k][0]) and torch.equal(x[pid][k][1], y[pid][k][1])
                    elif k == "policy_stats":
                        assert x[pid][k] == y[pid][k]
                    else:
                        assert torch.equal(x[pid][k], y[pid][k]), (pid, k)
    ca, cb = R.combine_batches([_fake_batch(seed=4), _fake_batch(seed=5)], cfg), combine_batches([_fake_batch(seed=4), _fake_batch(seed=5)], cfg)
    assert torch.equal(ca["obs"], cb["obs"]) and torch.equal(ca["policy_idxs"], cb["policy_idxs"])
