LS_EDGE_STAGED=0 timeout 300 python scripts/morton_experiment.py 2>&1 | tail -4
