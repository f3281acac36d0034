set -x
mkdir -p gpurun_out/r2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_edge|k_knn_tc|k_knn_rerank" -s 14 -c 9 -o gpurun_out/r2/prof_knn python scripts/one_forward.py 1 > gpurun_out/r2/prof_knn.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_fwd2.csv python scripts/one_forward.py 2 > gpurun_out/r2/launches_fwd2.log 2>&1
ls -la gpurun_out/r2/ | tail -5
