"""The planner's wave schedule (host arithmetic, no GPU): floors, growth, and the vectorised form the forest uses."""
import numpy as np


class _Sched:
    from posggym_baselines_b200.planning.mcts import MCTS as _M

    WAVE_FLOORS, wave_growth, max_wave = _M.WAVE_FLOORS, _M.WAVE_GROWTH, 262144
    _scheduled_wave, _scheduled_waves = _M._scheduled_wave, _M._scheduled_waves


def test_schedule_values():
    s = _Sched()
    # one warp on a young tree, two from 8 192 visits, then 1/2048 of the visits in whole warps, capped
    assert [s._scheduled_wave(v) for v in (0, 1, 8191)] == [32, 32, 32]
    assert [s._scheduled_wave(v) for v in (8192, 100_000, 131_072 + 65_535)] == [64, 64, 64]
    assert s._scheduled_wave(196_608) == 96 and s._scheduled_wave(1 << 20) == 512
    assert s._scheduled_wave(1 << 40) == s.max_wave
    # a wave never exceeds 1/2048 of the root's visits once past the floors
    for v in range(200_000, 5_000_000, 77_777):
        assert s._scheduled_wave(v) <= max(64, v // 2048)


def test_vectorised_schedule_equals_scalar():
    s = _Sched()
    rng = np.random.default_rng(0)
    v = np.concatenate([np.array([0, 1, 8191, 8192, 131071, 131072, 196607, 196608], np.int64),
                        rng.integers(0, 1 << 31, 2000)])
    assert list(s._scheduled_waves(v)) == [s._scheduled_wave(int(x)) for x in v]
