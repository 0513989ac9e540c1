# fused reduction in the tail of the epoch kernel vs the separate grid-wide reduce kernel (LCB_DECONV_REDUCE = auto | fused | separate), graph replays
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_deconv_gpu.py tests/test_api_gpu.py tests/test_starred_api_gpu.py -m gpu -q > gpurun_out/aa_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/aa_tests.log
for rep in 1; do
for mx in auto fused separate; do echo "== reduction route: $mx"; LCB_DECONV_REDUCE=$mx timeout 300 python tools/deconv_sweep.py --epochs 25,50,100,200 --cs 0 --iters 300; done
done 2>&1 | grep -v "^+" | tee gpurun_out/aa_ab.log
