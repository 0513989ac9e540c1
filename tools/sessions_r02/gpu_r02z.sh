set -x
mkdir -p gpurun_out
for rep in 1 2 3; do
echo "== units"; timeout 300 python tools/deconv_sweep.py --epochs 100,50 --cs 0 --iters 300
echo "== rows"; LCB_DC_FWDROWS=1 timeout 300 python tools/deconv_sweep.py --epochs 100,50 --cs 0 --iters 300
done 2>&1 | grep -v "^+" | tee gpurun_out/z_ab.log
