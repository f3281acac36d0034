mkdir -p gpurun_out/r2
timeout 600 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_bytes.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers --clock-control none -k regex:"k_edge_deep|k_knn_small" -c 12 --csv --log-file gpurun_out/r2/deep.csv python scripts/one_forward.py 2 > /dev/null 2>&1
python - <<PY
import csv
rows=[l for l in open("gpurun_out/r2/deep.csv") if not l.startswith("==")]
d={}
for r in csv.DictReader(rows):
    d.setdefault(r["ID"],{"n":r["Kernel Name"][:40],"g":r["Grid Size"]})[r["Metric Name"].split(".")[0][-28:]]=r["Metric Value"]
for k,v in d.items(): print(k,v)
PY
