ncu --set full --import-source on --clock-control none -k regex:search_waves -s 6 -c 1 -o gpurun_out/prof_search_waves_r2 python tools/waves_one.py 32768 > gpurun_out/ncu_waves.log 2>&1
ncu -i gpurun_out/prof_search_waves_r2.ncu-rep --page details 2>/dev/null | grep -E "Duration|Executed Ipc|Issue Slots|Registers Per|Local|Warp Cycles Per Issued|stall|Stall" | head -20
ncu -i gpurun_out/prof_search_waves_r2.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_search_waves_r2_src.csv 2>/dev/null
