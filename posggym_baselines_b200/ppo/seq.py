"""Sequence chunking of a rollout batch for BPTT on the device: the data format between the collector and the
PPO learner (/root/reference/posggym_baselines/ppo/utils.py:240-383 -- ``get_seq_idxs``,
``split_and_pad_batch``, ``split_lstm_state_batch`` walk every env in Python and build lists of index tensors; at
65 536 envs that alone takes seconds per update).

A chunk is ``(env, start, length)``: the reference's index tensor ``arange(start, start + length)`` of batch dim
``env``.  Chunks are ordered env-major, then by start, exactly like the reference's nested lists flattened.  Two
kernels (csrc/seq_kernels.cu): ``posg_seq_chunks`` finds every chunk in one launch (per-env scan of the dones
column, counts chained across CTAs); ``posg_seq_gather`` pads and gathers ALL entries of the batch in one launch."""
from __future__ import annotations

import ctypes as C
import struct
from typing import List, NamedTuple, Sequence, Tuple

import numpy as np
import torch

from posggym_baselines_b200 import _lib


class SeqChunks(NamedTuple):
    env: torch.Tensor      # [C] int32
    start: torch.Tensor    # [C] int32
    length: torch.Tensor   # [C] int32
    longest: int = 0       # longest chunk of the batch (the reference's pad_sequence length)

    def to_lists(self, num_envs: int) -> Tuple[List[List[torch.Tensor]], List[List[torch.Tensor]]]:
        """The reference's ``(seq_idxs, overlapping_seq_idxs)`` nested lists (host side; for tests and for code
        that still wants them)."""
        seq = [[] for _ in range(num_envs)]
        over = [[] for _ in range(num_envs)]
        for e, s, n in zip(self.env.tolist(), self.start.tolist(), self.length.tolist()):
            seq[e].append(torch.arange(s, s + n))
            over[e].append(torch.arange(s, s + n + 1))
        return seq, over


def get_seq_chunks(dones_buf: torch.Tensor, max_seq_len: int) -> SeqChunks:
    """``get_seq_idxs`` (ppo/utils.py:240-292): ``dones_buf`` ``[T, B]`` (non-zero where the step starts a new
    episode).  An episode piece starts at step 0 and at every done step; each piece is cut into chunks of at most
    ``max_seq_len`` steps.  One kernel launch and one read-back (the number of chunks sizes the outputs)."""
    T, B = dones_buf.shape
    dev = dones_buf.device
    d = dones_buf if dones_buf.dtype == torch.float32 and dones_buf.is_contiguous() else dones_buf.to(torch.float32).contiguous()
    cap = int(B) * int(T)     # every step its own chunk: the worst case
    env = torch.empty(cap, dtype=torch.int32, device=dev)
    start = torch.empty(cap, dtype=torch.int32, device=dev)
    length = torch.empty(cap, dtype=torch.int32, device=dev)
    res = np.zeros(2, np.int64)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().posg_seq_chunks(T, B, _lib.ptr(d), int(max_seq_len), _lib.ptr(env), _lib.ptr(start),
                                              _lib.ptr(length), cap, _lib.ptr(res), _lib.stream_ptr()))
    n = int(res[0])
    return SeqChunks(env[:n], start[:n], length[:n], int(res[1]))


def _pad_bits(value: float, dtype: torch.dtype) -> int:
    if dtype.is_floating_point:
        fmt = {2: None, 4: "<f", 8: "<d"}[torch.empty((), dtype=dtype).element_size()]
        if fmt is None:
            return int(torch.tensor(value, dtype=dtype).view(torch.int16).item()) & 0xFFFF
        return int.from_bytes(struct.pack(fmt, float(value)), "little")
    size = torch.empty((), dtype=dtype).element_size()
    return int(value) & ((1 << (8 * size)) - 1)


def split_and_pad_batches(batches: Sequence[torch.Tensor], chunks: SeqChunks, max_seq_len: int,
                          padding_values: Sequence[float] = (), overlapping: Sequence[bool] = ()) -> List[torch.Tensor]:
    """``split_and_pad_batch`` (ppo/utils.py:295-344) for every entry of the rollout batch in ONE launch: each
    ``batches[k]`` ``[T(+1), B, ...]`` -> ``[max_seq_len(+1), C, ...]``.  As in the reference (``pad_sequence``
    followed by ``F.pad``), rows past a chunk's end hold the padding value up to the longest chunk of the batch and
    0 after that.  ``overlapping[k]`` uses the reference's ``overlapping_seq_idxs`` (each chunk plus the first step
    after it; that entry then has ``T + 1`` rows)."""
    n = len(batches)
    pads = list(padding_values) + [0.0] * (n - len(padding_values))
    over = [bool(v) for v in overlapping] + [False] * (n - len(overlapping))
    Cn = int(chunks.env.shape[0])
    outs, done = [], 0
    while done < n:
        part = range(done, min(n, done + 16))
        src = (C.c_void_p * len(part))()
        dst = (C.c_void_p * len(part))()
        row_bytes = np.zeros(len(part), np.int64)
        ov = np.zeros(len(part), np.int32)
        eb = np.zeros(len(part), np.int32)
        pb = np.zeros(len(part), np.uint64)
        keep = []
        for j, k in enumerate(part):
            b = batches[k].contiguous()
            rows = max_seq_len + (1 if over[k] else 0)
            out = torch.empty((rows, Cn) + tuple(b.shape[2:]), dtype=b.dtype, device=b.device)
            keep.append(b)
            outs.append(out)
            src[j], dst[j] = b.data_ptr(), out.data_ptr()
            row_bytes[j] = b[0, 0].numel() * b.element_size() if b.dim() > 2 else b.element_size()
            ov[j], eb[j], pb[j] = int(over[k]), b.element_size(), _pad_bits(pads[k], b.dtype)
        if Cn:
            B = batches[part[0]].shape[1]
            with torch.cuda.device(chunks.env.device):
                _lib.check(_lib.lib().posg_seq_gather(
                    len(part), src, dst, _lib.ptr(row_bytes), _lib.ptr(ov), _lib.ptr(eb), _lib.ptr(pb), _lib.ptr(chunks.env),
                    _lib.ptr(chunks.start), _lib.ptr(chunks.length), Cn, B, int(max_seq_len), int(chunks.longest),
                    _lib.stream_ptr()))
        done += len(part)
    return outs


def split_and_pad_batch(batch: torch.Tensor, chunks: SeqChunks, max_seq_len: int, padding_value: float = 0.0,
                        overlapping: bool = False) -> torch.Tensor:
    """One entry (see ``split_and_pad_batches``)."""
    return split_and_pad_batches([batch], chunks, max_seq_len, [padding_value], [overlapping])[0]


def split_lstm_state_batch(lstm_state_batch: torch.Tensor, chunks: SeqChunks) -> torch.Tensor:
    """``split_lstm_state_batch`` (ppo/utils.py:347-383): ``[T, layers, B, N, H]`` -> the state at the first step
    of every chunk, ``[layers, C, N, H]``."""
    return lstm_state_batch[chunks.start.long(), :, chunks.env.long()].transpose(0, 1).contiguous()
