for v in "" _mb3 _mb4; do
echo "variant libbtgpu$v.so"
BTG_LIB=$PWD/bayestyper_b200/lib/libbtgpu$v.so BTG_NOISE_CONCURRENCY=1 BIGS=128 timeout 300 python tools/prof_real.py 0.33 2>&1 | grep -E "estimateNoise"
done
