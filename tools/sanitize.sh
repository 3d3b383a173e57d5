# compute-sanitizer over the small-N GPU tests that exercise the hand-rolled synchronisation:
# peer.cu / fused.cu release-acquire mailboxes (LocalPeerComm: shards on separate streams of one GPU),
# the cp.async.bulk + mbarrier ring of rank1_bulk_kernel, rankb_kernel, the potrf / trsm kernels.
# usage (under gpurun): bash tools/sanitize.sh ; summaries land in gpurun_out/sanitize_*.log
export TAL_SANITIZE=1
SAN=/usr/local/cuda/bin/compute-sanitizer
T1="tests/test_gpu_rankb.py::test_blocked_update_with_incremental_itl_and_general_m tests/test_gpu_kernels.py::test_rank1_variants_bit_identical tests/test_gpu_kernels.py::test_potrf_trsm_against_lapack"
T2="tests/test_gpu_fused.py::test_fused_step_row_sharded_peer tests/test_gpu_fused.py::test_failed_argmax_is_a_noop_on_the_device"
for tool in memcheck racecheck synccheck; do
  timeout 1500 $SAN --tool $tool --print-limit 20 python -m pytest $T1 -x -q -m gpu > gpurun_out/sanitize_${tool}_kernels.log 2>&1
  echo "$tool kernels rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_${tool}_kernels.log | tail -3
done
for tool in memcheck synccheck; do
  timeout 1500 $SAN --tool $tool --print-limit 20 python -m pytest $T2 -x -q -m gpu -k "2-1-lower or 2-2-full or noop" > gpurun_out/sanitize_${tool}_peer.log 2>&1
  echo "$tool peer rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_${tool}_peer.log | tail -3
done
