for cfg in "8 35" "8 25" "8 20" "16 25" "16 30" "16 20"; do
  set -- $cfg
  echo "sample=$1 x10=$2"
  HHFM_TOPN_SAMPLE=$1 HHFM_TOPN_CUT_X10=$2 python scripts/topn_stage_times.py --items 1000000 --reps 8 2>/dev/null
done
HHFM_TOPN_SAMPLE=8 HHFM_TOPN_CUT_X10=25 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -k "top or shard" 2>&1 | tail -3
HHFM_TOPN_SAMPLE=16 HHFM_TOPN_CUT_X10=25 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -k "top or shard" 2>&1 | tail -3
