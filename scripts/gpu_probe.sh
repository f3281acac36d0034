set -x
timeout 120 ./profiles/microbench/tc_probe_ts
