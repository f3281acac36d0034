set -x
mkdir -p gpurun_out/r2
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_fwd.csv python scripts/one_forward.py 2 > gpurun_out/r2/launches_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc2 -s 6 -c 4 -o gpurun_out/r2/prof_gemm2 python scripts/one_forward.py 1 > gpurun_out/r2/prof_gemm2.log 2>&1
ls -la gpurun_out/r2/
