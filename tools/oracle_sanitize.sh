#!/bin/bash
# Runs the CPU test suite against an AddressSanitizer + UBSan build of the oracle (oracle/Makefile target `san`) and of
# the host build of the witness kernels' per-thread logic (tools/hostsim.cpp).
# Usage: tools/oracle_sanitize.sh [pytest args]; log in /tmp/oracle_san.log, exit code = pytest's.
cd "$(dirname "$0")/.." || exit 1
make -C oracle -s san || exit 1
ASAN=$(gcc -print-file-name=libasan.so); UBSAN=$(gcc -print-file-name=libubsan.so)
TMX_HOSTSIM_SAN=1 TMX_ORACLE_LIB=$PWD/oracle/_build/liboracle_san.so LD_PRELOAD="$ASAN $UBSAN" ASAN_OPTIONS=detect_leaks=0 \
  UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 python -m pytest tests -x -q -m "not gpu" "$@" > /tmp/oracle_san.log 2>&1
rc=$?
grep -n "runtime error\|ERROR: AddressSanitizer" -A6 /tmp/oracle_san.log | cut -c1-200 | head -40
tail -3 /tmp/oracle_san.log
exit $rc
