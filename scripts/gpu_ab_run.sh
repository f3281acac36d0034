bash scripts/gpu_ab_variants.sh base stag160 stag320
