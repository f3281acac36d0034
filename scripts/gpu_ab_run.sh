bash scripts/gpu_ab_variants.sh s3 g1s4
