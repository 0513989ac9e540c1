set -x
mkdir -p gpurun_out
for mode in 0 1; do
LCB_DECONV_PUSH_LOCAL=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload deconv --steps 3 --warmup 2 --iters-per-step 400 --comm p2p > gpurun_out/bench_r02o_$mode.json 2> gpurun_out/bench_r02o_$mode.err; echo "rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_r02o_$mode.json') if l.startswith('{')][-1]; print('push_local=$mode', d['value'], {k:round(v['ms']/v['launches'],4) for k,v in d.get('kernels').items()})"
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-deconv > gpurun_out/bench_r02o_default.json 2> gpurun_out/bench_r02o_default.err; echo "default rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_r02o_default.json') if l.startswith('{')][-1]; print(d['value'], d['e2e']['value'], d['cpu_baseline'])"
