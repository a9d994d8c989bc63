set -x
for occ in 8 12; do
BTG_GIBBS_OCC=$occ timeout 300 python tools/prof_gibbs.py 30000 8 2>&1 | tail -2
BTG_GIBBS_OCC=$occ BTG_LIB=$PWD/bayestyper_b200/lib/libbtgpu_outline.so timeout 300 python tools/prof_gibbs.py 30000 8 2>&1 | tail -2
done
