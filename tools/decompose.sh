# Bottleneck decomposition of the tensor-core MLP kernels: run the per-kernel probe with the NERF_TC_DEBUG switches
# (mlp_tc.cuh) that remove one cost at a time.  Results with a switch set are numerically wrong by design.
#   1 no HBM stores   2 no MMAs   4 no weight copies   8 no epilogue   16 no dW drain   32 no dW bias sums
#   64 no chain launch   128 no dW launch
for f in ${FLAGS:-0 1 2 4 8 5 6 12 64 128 80 112 114}; do
  echo "== NERF_TC_DEBUG=$f"
  NERF_TC_DEBUG=$f timeout 120 python tools/kernel_probe.py --rays ${RAYS:-4096} --bwd --only mlp 2>&1 | grep -E "${ROWS:-M=  524288}|rror|rap"
done
