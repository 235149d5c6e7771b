set -x
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c5_st_launches.csv python scripts/bench_models.py --only fm_c5,hhfm_c5 --steps 2 > gpurun_out/c5_st_ncu.log 2>&1
HHFM_SINGLE_TOUCH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c5_nost_launches.csv python scripts/bench_models.py --only fm_c5,hhfm_c5 --steps 2 > gpurun_out/c5_nost_ncu.log 2>&1
tail -2 gpurun_out/c5_st_ncu.log
