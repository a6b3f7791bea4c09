nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/r2_pcie_ceiling.jsonl
python tools/pcie_ceiling.py >> gpurun_out/r2_pcie_ceiling.jsonl 2>> gpurun_out/r2_ceiling.err
for n in 2 4 8; do $TR --nproc-per-node $n --master-port 2951$n tools/pcie_ceiling.py >> gpurun_out/r2_pcie_ceiling.jsonl 2>> gpurun_out/r2_ceiling.err; done
for n in 2 4; do $TR --nproc-per-node $n --master-port 2952$n tools/pcie_ceiling.py --perm interleave >> gpurun_out/r2_pcie_ceiling.jsonl 2>> gpurun_out/r2_ceiling.err; done
cat gpurun_out/r2_pcie_ceiling.jsonl
: > gpurun_out/r2_c5_scale.jsonl
for n in 8 4; do $TR --nproc-per-node $n --master-port 2953$n bench.py --config c5 --gpus $n --steps 3 --warmup 3 >> gpurun_out/r2_c5_scale.jsonl 2>> gpurun_out/r2_c5_scale.err; done
cut -c1-600 gpurun_out/r2_c5_scale.jsonl
$TR --nproc-per-node 8 --master-port 29548 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_c3_8gpu.json 2> gpurun_out/r2_c3_8gpu.err
python -c "
import json
b=json.load(open('gpurun_out/r2_c3_8gpu.json'))
print('c3 N=8 value', b['value'], 'e2e', b['e2e'])
"
