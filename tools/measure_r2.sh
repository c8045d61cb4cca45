#!/bin/bash
# Round-2 measurement pass (run on the GPU box through gpurun): one `ncu --set full` capture per kernel the bench
# line reports, each after its driver program has run clean without ncu; the launch list of the bench command.
set -u
out=gpurun_out
cap() {  # name  kernel-regex  skip  command...
  name=$1; k=$2; skip=$3; shift 3
  "$@" > $out/${name}.run.log 2>&1 || { echo "$name: driver failed"; return; }
  ncu --set full --import-source on --clock-control none -k regex:"$k" -s $skip -c 1 -o $out/prof_${name} "$@" > $out/${name}.ncu.log 2>&1
  echo "$name rc=$?"
}
cap r2_step_fast_driving "step_fast_kernel" 5 python tools/step_one.py 16777216
cap r2_step_fast_pp "step_fast_kernel" 5 python tools/model_one.py pp 4194304
cap r2_step_fast_lbf "step_fast_kernel" 5 python tools/model_one.py lbf 4194304
cap r2_rollout_driving "rollout_kernel" 1 python tools/rollout_one.py 4194304
cap r2_search_waves "search_waves" 6 python tools/waves_one.py 32768
cap r2_policy_step "policy_step" 10 python tools/policy_one.py 8192
cap r2_vec_step_embed_8192 "vec_step_embed" 10 python tools/policy_one.py 8192
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --only-sweep > $out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
