"""Device-side pieces of the PPO worker / learner: the rollout collector with the policy on the tensor cores (SURVEY 8(f)
N1) and the batch post-processing (N2)."""
from posggym_baselines_b200.ppo.gae import compute_gae  # noqa: F401
from posggym_baselines_b200.ppo.batch import (combine_batches, filter_batch_by_partner,  # noqa: F401
                                              filter_policy_batches_by_partner_dist, split_batch_by_policy)
from posggym_baselines_b200.ppo.seq import (SeqChunks, get_seq_chunks, split_and_pad_batch, split_and_pad_batches,  # noqa: F401
                                            split_lstm_state_batch)
from posggym_baselines_b200.ppo.policy import FusedPolicy  # noqa: F401
from posggym_baselines_b200.ppo.collector import DeviceCollector  # noqa: F401
