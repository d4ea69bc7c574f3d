cd $GRAFT_REPO_ROOT
SNB_TUNE_VERBOSE=1 python scripts/gpu_tune.py 1024 text "5:1904" 2>&1 | grep -E "snb:|algo="
echo "L2_TABLES=0"; SNB_TUNE_L2_TABLES=0 python scripts/gpu_tune.py 1024 text "5:1904" 2>&1 | grep -E "algo="
for mb in 32 48 64; do echo "L2_MB=$mb"; SNB_TUNE_L2_MB=$mb python scripts/gpu_tune.py 1024 text "5:1904" 2>&1 | grep -E "algo="; done
for r in 0.5 0.75 1.0; do echo "ratio=$r"; SNB_TUNE_L2_RATIO=$r python scripts/gpu_tune.py 1024 text "5:1904" 2>&1 | grep -E "algo="; done
