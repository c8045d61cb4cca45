"""The collector's policy step on the tensor cores: ``PPOLSTMModel.get_action_and_value``
(/root/reference/posggym_baselines/ppo/network.py:212-268,305-330; called per rollout step at
ppo/worker.py:143-173) for the experiment network (trunk [64, 64], LSTM 64, heads [64]) as two launches --
``posg_vec_embed`` (trunk layer 0 on the packed observation) and ``posg_policy_step`` (everything else, one
tcgen05 kernel: csrc/policy_kernels.cu)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from posggym_baselines_b200 import _lib

_KEYS = ("trunk.2.weight", "trunk.2.bias", "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
         "actor.0.weight", "actor.0.bias", "actor.2.weight", "actor.2.bias", "critic.0.weight", "critic.0.bias",
         "critic.2.weight", "critic.2.bias")


class FusedPolicy:
    """Holds the packed weights of one ``PPOLSTMModel`` (``state_dict`` with the reference's parameter names)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: Optional[torch.device] = None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None:
            raise _lib.PosgB200Error("no CUDA device: posggym_baselines_b200 has no CPU fallback")
        self.device = torch.device(device)
        arrs = [np.ascontiguousarray(state_dict[k].detach().cpu().numpy(), dtype=np.float32) for k in _KEYS]
        self.hidden = int(arrs[0].shape[0])
        self.num_actions = int(arrs[8].shape[0])
        # trunk layer 0 (the embedding-sum of posg_vec_embed): transposed weight, bias
        self.w0_t = state_dict["trunk.0.weight"].detach().to(self.device, torch.float32).t().contiguous()
        self.b0 = state_dict["trunk.0.bias"].detach().to(self.device, torch.float32).contiguous()
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_policy_create(self.hidden, self.num_actions, *[_lib.ptr(a) for a in arrs],
                                                     C.byref(self._handle)))
            w0 = np.ascontiguousarray(state_dict["trunk.0.weight"].detach().cpu().numpy(), dtype=np.float32)
            b0 = np.ascontiguousarray(state_dict["trunk.0.bias"].detach().cpu().numpy(), dtype=np.float32)
            _lib.check(_lib.lib().posg_policy_set_trunk0(self._handle, int(w0.shape[1]), _lib.ptr(w0), _lib.ptr(b0)))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().posg_policy_destroy(h)
            except Exception:
                pass
            self._handle = None

    def eval_host(self, obs_index: np.ndarray, h: np.ndarray, c: np.ndarray):
        """The network for a few rows in host memory (``posg_policy_eval_host``; the planners' plugin path calls it once
        per node a simulation adds): ``obs_index`` [n, K] int32 = the active inputs of each one-hot-flattened
        observation, ``h`` / ``c`` [n, 64] float32.  Returns ``(h', c', logits [n, A], probs [n, A], value [n])`` as
        numpy arrays (``probs`` = softmax of the logits, computed on the device)."""
        obs_index = np.ascontiguousarray(obs_index, dtype=np.int32)
        n, k = obs_index.shape
        h = np.ascontiguousarray(h, dtype=np.float32).reshape(n, self.hidden)
        c = np.ascontiguousarray(c, dtype=np.float32).reshape(n, self.hidden)
        out = (np.empty((n, self.hidden), np.float32), np.empty((n, self.hidden), np.float32),
               np.empty((n, self.num_actions), np.float32), np.empty((n, self.num_actions), np.float32),
               np.empty((n,), np.float32))
        _lib.check(_lib.lib().posg_policy_eval_host(self._handle, n, k, _lib.ptr(obs_index), _lib.ptr(h), _lib.ptr(c),
                                                    *(_lib.ptr(o) for o in out)))
        return out

    def step(self, x0: torch.Tensor, lstm_state: Tuple[torch.Tensor, torch.Tensor], done: Optional[torch.Tensor],
             action: Optional[torch.Tensor] = None, seed: int = 0, step_idx: int = 0, row_offset: int = 0,
             want_logits: bool = False, out: Optional[dict] = None, terminated: Optional[torch.Tensor] = None,
             truncated: Optional[torch.Tensor] = None, step_device: Optional[torch.Tensor] = None):
        """``x0``: [n, 64] = tanh(trunk.0(obs)); ``lstm_state``: (h, c), each [n, 64] (updated IN PLACE);
        ``done``: [n] float, or None with ``terminated`` / ``truncated`` [n] uint8 (the vec env's buffers);
        ``step_device``: int32 device tensor whose first word is the Philox step index (captured graphs).
        Returns dict(action, logprob, entropy, value[, logits])."""
        n = x0.shape[0]
        h, c = lstm_state
        if out is None:
            out = {"action": torch.empty(n, dtype=torch.int64, device=self.device),
                   "logprob": torch.empty(n, dtype=torch.float32, device=self.device),
                   "entropy": torch.empty(n, dtype=torch.float32, device=self.device),
                   "value": torch.empty(n, dtype=torch.float32, device=self.device)}
            if want_logits:
                out["logits"] = torch.empty((n, self.num_actions), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_policy_step(
                self._handle, n, _lib.ptr(x0), _lib.ptr(h), _lib.ptr(c), _lib.ptr(done), _lib.ptr(terminated),
                _lib.ptr(truncated), seed, step_idx, _lib.ptr(step_device), row_offset,
                _lib.ptr(action), _lib.ptr(h), _lib.ptr(c), _lib.ptr(out["action"]), _lib.ptr(out["logprob"]),
                _lib.ptr(out["entropy"]), _lib.ptr(out["value"]), _lib.ptr(out.get("logits")), _lib.stream_ptr()))
        return out
