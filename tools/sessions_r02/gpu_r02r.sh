# round 2, session 3, call 2: per-epoch kernel with the warp-split forward pass / 16-byte push of r / merged block sums / prologue
# loads issued up front: parity (every deconvolution test + the PSF tests on the restored three-phase starlet), then timing
set -x
TAG=${TAG:-r}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest ${TESTS:-tests/test_deconv_gpu.py tests/test_api_gpu.py tests/test_starred_api_gpu.py} -m gpu -q --durations=8 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/${TAG}_tests.log
timeout 600 python tools/deconv_sweep.py --epochs 25,50,100,200 --cs 0 --iters 300 2>&1 | tee gpurun_out/${TAG}_sweep.log
timeout 600 python tools/deconv_sweep.py --epochs 25 --cs 4 --iters 300 2>&1 | tee -a gpurun_out/${TAG}_sweep.log
LCB_LIBRARY=lightcurver_b200/liblcb_dctim.so timeout 600 python tools/deconv_sweep.py --epochs 25,200 --cs 0 --iters 40 2>&1 | tee gpurun_out/${TAG}_dctim.log
