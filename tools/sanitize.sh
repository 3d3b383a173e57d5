# compute-sanitizer over the small-N GPU tests that exercise the hand-rolled synchronisation:
# peer.cu / fused.cu release-acquire mailboxes (LocalPeerComm: shards on separate streams of one GPU), the
# cp.async ring of rankb_async_kernel, the cp.async.bulk + mbarrier ring of rank1_bulk_kernel, the
# shared-memory DMMA kernels (potrf_diag_mma, gemm core of potrf / trsm).
# usage (under gpurun): bash tools/sanitize.sh ; summaries land in gpurun_out/sanitize_*.log
export TAL_SANITIZE=1
SAN=/usr/local/cuda/bin/compute-sanitizer
T1="tests/test_gpu_rankb.py::test_blocked_update_with_incremental_itl_and_general_m tests/test_gpu_kernels.py::test_rank1_variants_bit_identical tests/test_gpu_kernels.py::test_potrf_trsm_against_lapack"
T2="tests/test_gpu_fused.py::test_fused_step_row_sharded_peer tests/test_gpu_fused.py::test_failed_argmax_is_a_noop_on_the_device"
T3="tests/test_gpu_rankb.py::test_blocked_update_is_bit_identical"
timeout 600 $SAN --tool memcheck --print-limit 20 python -m pytest $T1 -x -q -m gpu > gpurun_out/sanitize_memcheck_kernels.log 2>&1
echo "memcheck kernels rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck_kernels.log | tail -3
timeout 600 $SAN --tool memcheck --print-limit 20 python -m pytest $T2 $T3 -x -q -m gpu -k "2-1-lower or 2-2-full or noop or (16 and lower and dtype0)" > gpurun_out/sanitize_memcheck_peer.log 2>&1
echo "memcheck peer rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck_peer.log | tail -3
timeout 900 $SAN --tool racecheck --print-limit 20 python -m pytest $T1 $T3 -x -q -m gpu -k "general_m or variants or potrf or (16 and lower and dtype0)" > gpurun_out/sanitize_racecheck_kernels.log 2>&1
echo "racecheck kernels rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitize_racecheck_kernels.log | tail -3
timeout 600 $SAN --tool synccheck --print-limit 20 python -m pytest $T1 -x -q -m gpu > gpurun_out/sanitize_synccheck_kernels.log 2>&1
echo "synccheck kernels rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_synccheck_kernels.log | tail -3
