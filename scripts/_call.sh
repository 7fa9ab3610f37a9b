bash scripts/gpu_round.sh
bash scripts/gpu_profile.sh
