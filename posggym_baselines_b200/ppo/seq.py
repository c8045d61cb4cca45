"""Sequence chunking of a rollout batch for BPTT, on the device and without a Python loop over
envs: the data format between the collector and the PPO learner
(/root/reference/posggym_baselines/ppo/utils.py:240-383 -- ``get_seq_idxs``,
``split_and_pad_batch``, ``split_lstm_state_batch`` walk every env in Python and build lists of
index tensors; at 65 536 envs that alone takes seconds per update).

A chunk is ``(env, start, length)``: the reference's index tensor ``arange(start, start + length)``
of batch dim ``env``.  Chunks are ordered env-major, then by start, exactly like the reference's
nested lists flattened."""
from __future__ import annotations

from typing import List, NamedTuple, Tuple

import torch


class SeqChunks(NamedTuple):
    env: torch.Tensor      # [C] int64
    start: torch.Tensor    # [C] int64
    length: torch.Tensor   # [C] int64

    def to_lists(self, num_envs: int) -> Tuple[List[List[torch.Tensor]], List[List[torch.Tensor]]]:
        """The reference's ``(seq_idxs, overlapping_seq_idxs)`` nested lists (host side; for tests
        and for code that still wants them)."""
        seq = [[] for _ in range(num_envs)]
        over = [[] for _ in range(num_envs)]
        for e, s, n in zip(self.env.tolist(), self.start.tolist(), self.length.tolist()):
            seq[e].append(torch.arange(s, s + n))
            over[e].append(torch.arange(s, s + n + 1))
        return seq, over


def get_seq_chunks(dones_buf: torch.Tensor, max_seq_len: int) -> SeqChunks:
    """``get_seq_idxs`` (ppo/utils.py:240-292): ``dones_buf`` ``[T, B]`` (non-zero where the step
    starts a new episode).  An episode piece starts at step 0 and at every done step; each piece is
    cut into chunks of at most ``max_seq_len`` steps."""
    T, B = dones_buf.shape
    dev = dones_buf.device
    t = torch.arange(T, device=dev).unsqueeze(1)                       # [T, 1]
    piece_start = dones_buf != 0
    piece_start = piece_start.clone()
    piece_start[0] = True
    # start step of the episode piece each step belongs to (running maximum of the start markers)
    start_of_piece = torch.cummax(torch.where(piece_start, t, torch.zeros_like(t)), dim=0).values
    chunk_start = piece_start | (((t - start_of_piece) % max_seq_len) == 0)
    es = torch.nonzero(chunk_start.t(), as_tuple=False)                # [C, 2] sorted by (env, step)
    env, start = es[:, 0].contiguous(), es[:, 1].contiguous()
    nxt = torch.full_like(start, T)
    if start.numel() > 1:
        same = env[1:] == env[:-1]
        nxt[:-1] = torch.where(same, start[1:], torch.full_like(start[1:], T))
    return SeqChunks(env, start, nxt - start)


def split_and_pad_batch(batch: torch.Tensor, chunks: SeqChunks, max_seq_len: int, padding_value: float = 0.0,
                        overlapping: bool = False) -> torch.Tensor:
    """``split_and_pad_batch`` (ppo/utils.py:295-344): ``batch`` ``[T(+1), B, ...]`` ->
    ``[max_seq_len(+1), C, ...]``.  As in the reference (``pad_sequence`` followed by ``F.pad``), rows
    past a chunk's end hold ``padding_value`` up to the longest chunk of the batch and 0 after that.
    ``overlapping=True`` uses the reference's ``overlapping_seq_idxs`` (each chunk plus the first
    step after it; ``batch`` then has ``T + 1`` rows)."""
    extra = 1 if overlapping else 0
    rows = max_seq_len + extra
    length = chunks.length + extra
    s = torch.arange(rows, device=batch.device).unsqueeze(1)           # [rows, 1]
    valid = s < length.unsqueeze(0)                                    # [rows, C]
    tidx = (chunks.start.unsqueeze(0) + s).clamp(max=batch.shape[0] - 1)
    out = batch[tidx, chunks.env.unsqueeze(0)]                         # [rows, C, ...]
    longest = length.max() if length.numel() else torch.zeros((), dtype=torch.int64, device=batch.device)
    pad = torch.where(s < longest, torch.full((), padding_value, device=batch.device, dtype=batch.dtype),
                      torch.zeros((), device=batch.device, dtype=batch.dtype))          # [rows, 1]
    shape = valid.shape + (1,) * (batch.dim() - 2)
    return torch.where(valid.reshape(shape), out, pad.reshape((rows, 1) + (1,) * (batch.dim() - 2)))


def split_lstm_state_batch(lstm_state_batch: torch.Tensor, chunks: SeqChunks) -> torch.Tensor:
    """``split_lstm_state_batch`` (ppo/utils.py:347-383): ``[T, layers, B, N, H]`` -> the state at the
    first step of every chunk, ``[layers, C, N, H]``."""
    return lstm_state_batch[chunks.start, :, chunks.env].transpose(0, 1).contiguous()
