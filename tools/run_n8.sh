TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29511 tools/pcie_probe.py > gpurun_out/r2_pcie_n8_unbound.json 2> gpurun_out/r2_pcie_n8.err
$TR --master-port 29512 tools/pcie_probe.py --bind > gpurun_out/r2_pcie_n8_bound.json 2>> gpurun_out/r2_pcie_n8.err
python tools/pcie_probe.py > gpurun_out/r2_pcie_n1.json 2>> gpurun_out/r2_pcie_n8.err
cat gpurun_out/r2_pcie_n8_unbound.json gpurun_out/r2_pcie_n8_bound.json gpurun_out/r2_pcie_n1.json
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" > gpurun_out/r2_lscpu.txt
$TR --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n8.json')); print(d['value'], d['e2e']['value'], d['e2e']['compact_results'].get('value'), d['e2e']['host_binding']); print(d['pomcp']['predator_prey_root_parallel']); print(d['collector'])"
