#!/bin/bash
# compute-sanitizer over the GPU test-suite (run on the GPU box: gpurun -- 'bash tools/sanitize.sh memcheck tests/...').
#   --report-api-errors no : the CUDA runtime's lazy-loading probe (cuKernelGetFunction -> CUDA_ERROR_INVALID_HANDLE on the first
#                            launch of every process) is an API return code, not a memory error (profiles/r02_sanitizer.md)
#   deselected: the test that itself spawns compute-sanitizer, and CUDA-graph capture (impossible without the caching allocator)
#   PYTORCH_NO_CUDA_MEMORY_CACHING=1 : every tensor is its own cudaMalloc, so an out-of-bounds access cannot land inside a pool
tool=${1:-memcheck}; shift
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout ${SANITIZE_TIMEOUT:-1500} compute-sanitizer --tool $tool --report-api-errors no \
    --print-limit 60 --log-file gpurun_out/sanitize_$tool.log python -m pytest "$@" -q -m gpu -p no:cacheprovider \
    --deselect tests/test_conv_gpu.py::test_fused_epilogue_multi_image_tiles_under_memcheck \
    --deselect tests/test_step_gpu.py::test_cuda_graph_replay_matches_eager \
    > gpurun_out/sanitize_${tool}_pytest.log 2>&1
echo "== $tool: pytest rc=$?"; tail -3 gpurun_out/sanitize_${tool}_pytest.log; grep -c "=========" gpurun_out/sanitize_$tool.log; tail -4 gpurun_out/sanitize_$tool.log
