set -x
mkdir -p gpurun_out/r2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc2 -s 9 -c 2 -o gpurun_out/r2/prof_sdf_gemm python scripts/one_sdf.py 2 > gpurun_out/r2/prof_sdf_gemm.log 2>&1
tail -3 gpurun_out/r2/prof_sdf_gemm.log
ls -la gpurun_out/r2/*.ncu-rep
