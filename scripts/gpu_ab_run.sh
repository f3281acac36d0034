bash scripts/gpu_ab_variants.sh rr3 rr4
