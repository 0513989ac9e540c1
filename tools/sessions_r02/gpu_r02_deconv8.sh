N=$1
set -x
mkdir -p gpurun_out
for comm in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload deconv --steps 3 --warmup 2 --iters-per-step 400 --comm $comm > gpurun_out/bench_r02b_deconv_${N}gpu_$comm.json 2> gpurun_out/bench_r02b_deconv_${N}gpu_$comm.err; echo "deconv $comm rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_r02b_deconv_${N}gpu_$comm.json') if l.startswith('{')][-1]; print('deconv $comm', d['value'], d.get('kernels'))"
done
