set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deconv_gpu.py tests/test_psf_gpu.py tests/test_api_gpu.py tests/test_starred_api_gpu.py -m gpu -q -s -p no:cacheprovider -k "not configured_length" > gpurun_out/tests_r02h.log 2>&1; echo "tests rc=$?"; grep -v "^\[parity\]" gpurun_out/tests_r02h.log | tail -8
timeout 600 python bench.py --workload deconv --steps 3 --warmup 2 --iters-per-step 400 --no-cpu-baseline > gpurun_out/bench_r02h_deconv_1gpu.json 2> gpurun_out/bench_r02h_deconv_1gpu.err; echo "rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_r02h_deconv_1gpu.json') if l.startswith('{')][-1]; print('deconv 1gpu', d['value'], d.get('kernels'))"
LCB_DECONV_STARLET_CLUSTER=1 timeout 600 python bench.py --workload deconv --steps 3 --warmup 2 --iters-per-step 400 --no-cpu-baseline > gpurun_out/bench_r02h_deconv_1gpu_cl.json 2>/dev/null
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_r02h_deconv_1gpu_cl.json') if l.startswith('{')][-1]; print('deconv 1gpu (cluster starlet)', d['value'], d.get('kernels'))"
