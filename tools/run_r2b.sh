python -m pytest tests/test_pomcp_gpu.py tests/test_intmcp_gpu.py tests/test_belief_gpu.py tests/test_pp_gpu.py tests/test_lbf_gpu.py -x -q --durations=8 2>&1 | tail -25
python tools/update_probe.py 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo bench rc=$?; tail -3 gpurun_out/r2b_bench_n1.err
