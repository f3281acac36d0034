# Round-2 evidence run (1 x B200): full GPU test suite, smoke(), bench lines (both arms), ncu launch lists and
# --set full captures of the top kernels.  Outputs under gpurun_out/r2final/ (summaries are copied to profiles/r02/).
set -x
O=gpurun_out/r2final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/pytest_gpu.log; cat $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 > $O/smoke.log; cat $O/smoke.log
timeout 900 python bench.py > $O/bench_r02.json 2> $O/bench_r02.err; tail -2 $O/bench_r02.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_r02.json 2> $O/bench_ref_r02.err; tail -2 $O/bench_ref_r02.err
timeout 600 python bench.py --workload sdf > $O/bench_sdf_r02.json 2> $O/bench_sdf_r02.err
# launch list of the bench command itself (per-launch times are serialised, cold-cache: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-c4 --no-sdf > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
# one encoder forward with DRAM / tensor / L2 metrics per launch
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -c 400 --csv --log-file $O/launches_fwd.csv python scripts/one_forward.py 2 > $O/launches_fwd.log 2>&1
# --set full captures (second forward): GEMMs, kNN path, EdgeConv; SDF GEMM
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc" -s 18 -c 18 -o $O/prof_gemm python scripts/one_forward.py 2 > $O/prof_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc|k_knn_rerank|k_knn_edge" -s 17 -c 17 -o $O/prof_knn python scripts/one_forward.py 2 > $O/prof_knn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc3" -s 8 -c 3 -o $O/prof_sdf python scripts/one_sdf.py 2 > $O/prof_sdf.log 2>&1
# gpurun merges at most 64 MiB back: keep the raw metric tables (csv) of every capture and the .ncu-rep of the SDF GEMM only
for r in prof_gemm prof_knn prof_sdf; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null; done
ncu -i $O/prof_knn.ncu-rep --page source --csv --kernel-id :::5 > $O/prof_knn_tc_layer1.source.csv 2>/dev/null
ncu -i $O/prof_knn.ncu-rep --page source --csv --kernel-id :::9 > $O/prof_knn_edge_layer2.source.csv 2>/dev/null
rm -f $O/prof_gemm.ncu-rep $O/prof_knn.ncu-rep
ls -la $O
du -sh $O
