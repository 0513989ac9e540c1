set -x
mkdir -p gpurun_out
timeout 600 python tools/deconv_sweep.py --iters 100 --epochs 25,50,100,200 --cs 0 > gpurun_out/sweep_r02l.log 2>&1; cat gpurun_out/sweep_r02l.log
