bash scripts/gpu_ab_variants.sh base rr3 rr4 kt1
