cd $GRAFT_REPO_ROOT
for v in 16 32 64; do
SPLENDOR_B200_LIB=$GRAFT_REPO_ROOT/gpurun_variants/lib_wb$v.so QUIET=1 timeout 300 python tools/explore.py --beam 30000000 --reps 3 2>&1 | grep -E "SUMMARY" | tail -1
done
