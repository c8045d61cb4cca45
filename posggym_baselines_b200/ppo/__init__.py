"""Device-side pieces of the PPO learner's batch post-processing."""
from posggym_baselines_b200.ppo.gae import compute_gae  # noqa: F401
from posggym_baselines_b200.ppo.seq import SeqChunks, get_seq_chunks, split_and_pad_batch, split_lstm_state_batch  # noqa: F401
from posggym_baselines_b200.ppo.policy import FusedPolicy  # noqa: F401
