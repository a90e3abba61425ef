for f in test_gpu_sddmm test_gpu_layers test_gpu_fullsize test_gpu_vs_reference test_gpu_sharding; do
  echo "=== $f"; timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 600 -x 2>&1 | grep -v Warn | tail -2
done
for wl in reddit-like-rmat reddit-like-uniform products-like-rmat rmat-10m-200m citeseer-like; do
  timeout 300 python tools/quick.py --workload $wl --op sddmm --iters 3 2>&1 | tail -1
done
timeout 300 python tools/quick.py --workload reddit-like-uniform --op sddmm --dim 32 --iters 3 2>&1 | tail -1
timeout 300 python tools/quick.py --workload reddit-like-uniform --op sddmm --dim 96 --iters 3 2>&1 | tail -1
