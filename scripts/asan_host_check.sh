#!/bin/bash
# Host control code under AddressSanitizer + UBSan: builds the host-logic check library (host C
# objects + the CPU restatement of the kernels) with -fsanitize=address,undefined and runs the CPU
# parity tests that drive it.  Usage: bash scripts/asan_host_check.sh   (no GPU needed)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OB=${OB:-/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs}
TMP=$(mktemp -d)
gcc -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -std=c99 -D_GNU_SOURCE -shared \
   -o $TMP/libprimme_hostcheck.so $ROOT/primme_b200/src/*.c $ROOT/oracle/kernels_ref.c \
   -L$OB -l:libopenblasp-r0-59ffcd50.3.15.so -Wl,--disable-new-dtags,-rpath,$OB -lm
cp $ROOT/oracle/_build/libprimme_hostcheck.so $TMP/orig.so
trap "cp $TMP/orig.so $ROOT/oracle/_build/libprimme_hostcheck.so" EXIT
cp $TMP/libprimme_hostcheck.so $ROOT/oracle/_build/libprimme_hostcheck.so
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 \
   python -m pytest $ROOT/tests/test_jdqmr_cpu.py $ROOT/tests/test_svds_cpu.py $ROOT/tests/test_host_logic.py \
   $ROOT/tests/test_reference_golden_cpu.py $ROOT/tests/test_driver_cpu.py $ROOT/tests/test_refined_cpu.py $ROOT/tests/test_edge_cases_cpu.py $ROOT/tests/test_python_api.py -q -m "not gpu" 2>&1 | tee $TMP/out.txt | tail -3
echo "sanitizer reports: $(grep -c 'AddressSanitizer\|runtime error' $TMP/out.txt || true)"
