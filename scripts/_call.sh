bash scripts/gpu_round.sh
