mkdir -p gpurun_out
KFACB_PROFILE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 2>&1 | grep -E "kfacb eigh profile|metric" | cut -c1-300
