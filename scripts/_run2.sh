export PB200_DEBUG=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "ortho" --timeout 120 2>&1 | tail -5
echo "=== kernel bench (exact)"
timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | grep "ortho\|projection"
echo "=== kernel bench (generic)"
PB200_NO_ORTHO_EXACT=1 timeout 200 python scripts/kernel_bench.py --reps 10 2>&1 | grep "ortho\|projection"
