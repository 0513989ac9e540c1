# two GPUs: the two-rank tests on both reduction routes, then the sharded cfg4 bench on the default (separate) and the fused route
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_deconv_gpu.py -m gpu -q -p no:cacheprovider -k "two_ranks" > gpurun_out/ad_tests.log 2>&1; echo "two-rank tests rc=$?"; tail -2 gpurun_out/ad_tests.log
LCB_DECONV_REDUCE=fused timeout 600 python -m pytest tests/test_deconv_gpu.py -m gpu -q -p no:cacheprovider -k "two_ranks and p2p" > gpurun_out/ad_tests_fused.log 2>&1; echo "two-rank tests (fused) rc=$?"; tail -2 gpurun_out/ad_tests_fused.log
for mx in auto fused; do
LCB_DECONV_REDUCE=$mx timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload deconv --steps 3 --warmup 2 --iters-per-step 400 --comm p2p --no-cpu-baseline > gpurun_out/bench_r02d_deconv_2gpu_p2p_$mx.json 2> gpurun_out/bench_r02d_deconv_2gpu_p2p_$mx.err; echo "deconv $mx rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_r02d_deconv_2gpu_p2p_$mx.json') if l.startswith('{')][-1]; print('$mx', d['value'], {k: round(v['ms']/v['launches'],4) for k,v in d['kernels'].items()})"
done
