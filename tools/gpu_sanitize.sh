#!/bin/bash
# compute-sanitizer over small proofs and the kernel-level entry points (run under gpurun; logs in gpurun_out/).
#   memcheck  : out-of-bounds / misaligned global, shared and local accesses, CUDA API errors
#   racecheck : shared-memory hazards (NTT tiles, quotient staging, Merkle levels)
#   initcheck : reads of uninitialised device memory
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
PROOFS='tests/test_gpu_prove.py::test_proof_bytes_equal_oracle tests/test_gpu_witness.py::test_witness_tables_match_oracle_on_fixtures'
PRIMS='tests/test_gpu_primitives.py'
run() {  # name tool timeout tests...
  local name=$1 tool=$2 t=$3; shift 3
  timeout "$t" $SAN --tool "$tool" --error-exitcode 86 --print-limit 20 --log-file gpurun_out/san_${name}.log \
      python -m pytest -x -q -m gpu "$@" > gpurun_out/san_${name}.pytest.log 2>&1
  echo "$name: exit $? | $(tail -1 gpurun_out/san_${name}.pytest.log) | $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/san_${name}.log | tail -1)"
}
run memcheck_proofs memcheck 420 $PROOFS
run memcheck_prims memcheck 300 $PRIMS -k "not full_size and not 17 and not 16"
run racecheck_proofs racecheck 420 tests/test_gpu_prove.py::test_proof_bytes_equal_oracle -k "skip_3000_3100_n4"
run initcheck_proofs initcheck 300 tests/test_gpu_prove.py::test_proof_bytes_equal_oracle -k "skip_3000_3100_n4 or step_10500"
run racecheck_prims racecheck 420 $PRIMS -k "not full_size"
run synccheck_proofs synccheck 300 tests/test_gpu_prove.py::test_proof_bytes_equal_oracle -k "skip_3000_3100_n4 or step_10500"
run memcheck_ragged memcheck 600 tests/test_gpu_prove.py::test_ragged_validator_sets_equal_oracle tests/test_gpu_prove.py::test_unsat_reports_the_failing_check tests/test_gpu_witness.py
run memcheck_prims_large memcheck 420 $PRIMS -k "17 or 16"
