# N-GPU A/B of the reduction route of the sharded joint deconvolution (LCB_DECONV_REDUCE = separate: the reduce kernel pushes
# the sums; fused: the tail of the epoch kernel pushes them), then the default bench line with the faster route
N=${1:-2}; tag=${2:-r02w}
set -x
mkdir -p gpurun_out
best=auto; bestv=0
for mx in separate fused; do
LCB_DECONV_REDUCE=$mx timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload deconv --steps 3 --warmup 2 --iters-per-step 400 --comm p2p --no-cpu-baseline > gpurun_out/bench_${tag}_deconv_${N}gpu_p2p_fused$mx.json 2> gpurun_out/bench_${tag}_deconv_${N}gpu_p2p_fused$mx.err; echo "deconv fused_max=$mx rc=$?"
v=$(python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_${tag}_deconv_${N}gpu_p2p_fused$mx.json') if l.startswith('{')][-1]; print(int(d['value']))")
echo "fused_max=$mx -> $v it/s"
if [ "$v" -gt "$bestv" ]; then bestv=$v; best=$mx; fi
done
echo "best route: LCB_DECONV_REDUCE=$best ($bestv it/s)"
if [ "$3" = "bench" ]; then
LCB_DECONV_REDUCE=$best timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${tag}_${N}gpu.json 2> gpurun_out/bench_${tag}_${N}gpu.err; echo "bench rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_${tag}_${N}gpu.json') if l.startswith('{')][-1]; print('psfphot', d['value'], 'e2e', d['e2e']['value'], 'deconv', d['deconv']['value'], d['deconv'].get('parity_vs_single_rank',{}).get('ok'), d['deconv']['kernels'])"
fi
