"""Sequence chunking of rollout batches (posggym_baselines_b200.ppo.seq, vectorised torch) against
the per-env loops of the reference (restated in oracle/ppo_utils.py and, when /root/reference is
importable, the reference module itself).  Runs on the CPU; the same code runs on CUDA tensors."""
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


@pytest.mark.parametrize("dones,L", cases())
def test_chunks_and_padding_match_reference_loops(dones, L):
    T, B = dones.shape
    ref_idx, ref_over = O.get_seq_idxs(dones, L)
    ch = get_seq_chunks(dones, L)
    got_idx, got_over = ch.to_lists(B)
    assert [[c.tolist() for c in e] for e in got_idx] == [[c.tolist() for c in e] for e in ref_idx]
    assert [[c.tolist() for c in e] for e in got_over] == [[c.tolist() for c in e] for e in ref_over]
    g = torch.Generator().manual_seed(1)
    batch = torch.randn((T, B, 2, 3), generator=g)
    for pad in (0.0, -1.5):
        assert torch.equal(split_and_pad_batch(batch, ch, L, pad), O.split_and_pad_batch(batch, ref_idx, L, pad))
    batch1 = torch.randn((T + 1, B, 2), generator=g)
    assert torch.equal(split_and_pad_batch(batch1, ch, L, 0.0, overlapping=True),
                       O.split_and_pad_batch(batch1, ref_over, L + 1, 0.0))
    ints = torch.randint(0, 6, (T, B, 2), generator=g)
    assert torch.equal(split_and_pad_batch(ints, ch, L, 0.0), O.split_and_pad_batch(ints, ref_idx, L, 0.0))
    lstm = torch.randn((T, 1, B, 2, 4), generator=g)
    assert torch.equal(split_lstm_state_batch(lstm, ch), O.split_lstm_state_batch(lstm, ref_idx))


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
    ch_cpu = get_seq_chunks(dones, 16)
    for a, b in zip(ch_gpu, ch_cpu):
        assert torch.equal(a.cpu(), b)
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
