python -m pytest tests/test_pomcp_gpu.py tests/test_intmcp_gpu.py tests/test_pp_gpu.py tests/test_lbf_gpu.py -q --durations=8 2>&1 | tail -30
python tools/episode_probe.py 0.1 8 2>&1 | tail -9
