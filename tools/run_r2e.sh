ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"policy_step|vec_embed|vec_step" -s 600 -c 12 --csv --log-file gpurun_out/r2e_collector_launches.csv python tools/collector_probe.py > gpurun_out/r2e_collector.log 2>&1
grep -v "^==" gpurun_out/r2e_collector_launches.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | head -14
