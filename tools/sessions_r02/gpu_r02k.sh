set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deconv_gpu.py tests/test_api_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/tests_r02k.log 2>&1; echo "tests rc=$?"; grep -v "^\[parity\]" gpurun_out/tests_r02k.log | tail -5
for comm in p2p; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload deconv --steps 3 --warmup 2 --iters-per-step 400 --comm $comm > gpurun_out/bench_r02c_deconv_2gpu_$comm.json 2> gpurun_out/bench_r02c_deconv_2gpu_$comm.err; echo "deconv $comm rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_r02c_deconv_2gpu_$comm.json') if l.startswith('{')][-1]; print('deconv $comm', d['value'], d.get('kernels'), d.get('parity_vs_single_rank'))"
done
