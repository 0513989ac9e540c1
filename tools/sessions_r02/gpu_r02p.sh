set -x
mkdir -p gpurun_out
for comm in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload deconv --steps 3 --warmup 2 --iters-per-step 400 --comm $comm > gpurun_out/bench_r02p_$comm.json 2> gpurun_out/bench_r02p_$comm.err; echo "rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_r02p_$comm.json') if l.startswith('{')][-1]; print('$comm', d['value'], d.get('kernel_ms_per_launch_by_rank'))"
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,pci.bus_id --format=csv
nvidia-smi topo -m | head -12
