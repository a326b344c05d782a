bash scripts/gpu_scale.sh $1 r2
