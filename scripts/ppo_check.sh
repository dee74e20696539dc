python -m pytest tests/test_gpu_rl.py -x -q 2>&1 | tail -8
python -m rui_b200.rl --config examples/rl_config_smoke.yaml --num-envs 8192 --n-steps 32 --total-timesteps 4e6 2>&1 | grep "ppo\] it" | tail -4 | cut -c1-120
