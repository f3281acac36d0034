set -x
mkdir -p gpurun_out/r2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc|k_knn_rerank|k_knn_pack" -s 17 -c 17 -o gpurun_out/r2/prof_knn_split python scripts/one_forward.py 2 > gpurun_out/r2/prof_knn_split.log 2>&1
tail -2 gpurun_out/r2/prof_knn_split.log
ls -la gpurun_out/r2/*.ncu-rep
