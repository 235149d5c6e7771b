for cfg in "8 35" "16 35" "16 40" "16 45" "16 50" "32 50"; do
  set -- $cfg
  echo "sample=$1 x10=$2"
  HHFM_TOPN_SAMPLE=$1 HHFM_TOPN_CUT_X10=$2 python scripts/topn_stage_times.py --contexts 65536 --items 1000000 --reps 4 2>/dev/null
done
