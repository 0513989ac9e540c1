# two-GPU validation of the sharded joint deconvolution after the per-epoch kernel changes of this session
N=${1:-2}; tag=${2:-r02u}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
[ "$N" -le 2 ] && timeout 900 python -m pytest tests/test_deconv_gpu.py tests/test_api_gpu.py -m gpu -q -s -p no:cacheprovider -k "two_ranks or multi_gpu or fan_out" > gpurun_out/tests_multi_$tag.log 2>&1; echo "multi tests rc=$?"; tail -5 gpurun_out/tests_multi_$tag.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${tag}_${N}gpu.json 2> gpurun_out/bench_${tag}_${N}gpu.err; echo "bench rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_${tag}_${N}gpu.json') if l.startswith('{')][-1]; print('psfphot', d['value'], 'e2e', d['e2e']['value'], 'deconv', d['deconv']['value'], d['deconv'].get('parity_vs_single_rank'), d['deconv']['kernels'])"
for comm in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload deconv --steps 3 --warmup 2 --iters-per-step 400 --comm $comm > gpurun_out/bench_${tag}_deconv_${N}gpu_$comm.json 2> gpurun_out/bench_${tag}_deconv_${N}gpu_$comm.err; echo "deconv $comm rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_${tag}_deconv_${N}gpu_$comm.json') if l.startswith('{')][-1]; print('deconv $comm', d['value'], d.get('kernels'))"
done
