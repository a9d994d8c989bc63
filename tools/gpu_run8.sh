set -x
mkdir -p gpurun_out
BIGS=128 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_noise_chain -c 1 -o gpurun_out/r1i_noise_real -f python tools/prof_real.py 0.33 > gpurun_out/r1i_ncu_noise_real.log 2>&1
tail -5 gpurun_out/r1i_ncu_noise_real.log
