#!/bin/bash
# compute-sanitizer over the small parity cases (run on the B200 box: bash tools/gpu.sh sanitize).
# memcheck: out-of-bounds / misaligned accesses of every kernel the selected tests launch;
# racecheck: shared-memory hazards (the tcgen05 kernel's mbarrier pipeline, the shared-memory
# tables of the categorical kernels).  10-100x slower than a plain run, hence the small selection.
set -u
mkdir -p gpurun_out
SEL=${SEL:-"tests/test_gpu_fused_scatter.py::test_fused_scatter_counts_are_exact tests/test_gpu_dense_tc.py::test_tcgen05_syrk_row_restriction tests/test_gpu_index_fused.py::test_fused_index_counts_are_exact tests/test_gpu_boundary_extras.py tests/test_gpu_irls.py::test_default_implementation_for_leaf_matrices tests/test_gpu_index_fused.py::test_row_blocked_csc_path tests/test_gpu_dense_pad.py::test_split_matrix_with_padded_dense_block"}
for tool in memcheck racecheck; do
  timeout -s KILL ${T:-1200} compute-sanitizer --tool $tool --error-exitcode 99 --log-file gpurun_out/sanitize_$tool.log \
    python -m pytest $SEL -m gpu -q -x -p no:cacheprovider > gpurun_out/sanitize_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitize_${tool}_pytest.log; grep -E "ERROR SUMMARY|Error:|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -5
done
