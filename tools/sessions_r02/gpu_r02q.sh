# round 2, session 3, call 1: parity of the two-phase starlet adjoint (K1 + K3), A/B against the three-phase one, cluster-size
# sweep of the per-epoch kernel at the 8-GPU shard size (25 local epochs) with the phase stamps of the -DLCB_DC_TIMERS build
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1200 python -m pytest tests/test_psf_gpu.py tests/test_deconv_gpu.py tests/test_golden_gpu.py -m gpu -x -q --durations=8 > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?"; tail -20 gpurun_out/q_tests.log
for rep in 1 2; do
for v in "" _v1; do echo "lib$v"; AB_F=592 AB_T=300 LCB_LIBRARY=lightcurver_b200/liblcb$v.so timeout 300 python tools/ab_time.py; done
done 2>&1 | tee gpurun_out/q_ab.log
timeout 600 python tools/deconv_sweep.py --epochs 25 --cs 3,4,5,6,7,8 --iters 300 2>&1 | tee gpurun_out/q_sweep.log
timeout 600 python tools/deconv_sweep.py --epochs 50,100,200 --cs 0,3,5,6 --iters 100 2>&1 | tee -a gpurun_out/q_sweep.log
LCB_LIBRARY=lightcurver_b200/liblcb_v1.so timeout 300 python tools/deconv_sweep.py --epochs 25 --cs 8 --iters 300 2>&1 | tee -a gpurun_out/q_sweep.log
LCB_DECONV_GRAPH=0 LCB_LIBRARY=lightcurver_b200/liblcb_dctim.so timeout 600 python tools/deconv_sweep.py --epochs 25,200 --cs 0,6 --iters 40 2>&1 | tee gpurun_out/q_dctim.log
