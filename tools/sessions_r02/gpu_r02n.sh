N=$1
set -x
mkdir -p gpurun_out
time (timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err); echo "bench rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_r02_${N}gpu.json') if l.startswith('{')][-1]; print('psfphot', d['value'], 'e2e', d['e2e']['value'], 'deconv', d['deconv']['value'], d['deconv'].get('parity_vs_single_rank',{}).get('ok'), d['deconv']['kernels'])"
time (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/benchref_r02_${N}gpu.json 2> gpurun_out/benchref_r02_${N}gpu.err); echo "ref rc=$?"; tail -c 300 gpurun_out/benchref_r02_${N}gpu.json
