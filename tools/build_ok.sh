#!/bin/bash
# Build everything; exit non-zero (and say so loudly) if anything failed.  Run before every gpurun.
cd "$(dirname "$0")/../query-compiler-executor_b200" || exit 1
if ! make all > /tmp/qce_build.log 2>&1; then
    grep -E "error" /tmp/qce_build.log build_ptxas.log 2>/dev/null | head -10
    echo "BUILD FAILED"; exit 1
fi
if grep -qE "error" build_ptxas.log 2>/dev/null; then grep -E "error" build_ptxas.log | head; echo "BUILD FAILED"; exit 1; fi
echo "build ok: $(ls -la --time-style=+%T libqce_b200.so | awk '{print $6}') now $(date +%T)"
