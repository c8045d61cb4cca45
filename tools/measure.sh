#!/bin/bash
# Round measurement pass (run on the GPU box through gpurun): bench line, ncu launch list of the
# same command, one full ncu capture of the dominant kernel at the bench's batch size.
# usage: tools/measure.sh <tag>      outputs under gpurun_out/
set -u
tag=${1:-r1}
out=gpurun_out
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
echo "reference rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:step_fast -s 5 -c 1 -o $out/prof_step_${tag} \
    python tools/step_one.py 16777216 > $out/${tag}_ncu_step.log 2>&1
echo "ncu step rc=$?"
