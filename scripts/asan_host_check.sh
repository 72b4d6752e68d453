#!/bin/bash
# Host control code under AddressSanitizer + UBSan: builds the host-logic check library (host C objects, both the real
# and the -DPB_COMPLEX instantiation of the typed sources, + the CPU restatement of the kernels) with
# -fsanitize=address,undefined and runs the CPU parity tests that drive it.  Usage: bash scripts/asan_host_check.sh [tests...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OB=${OB:-/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs}
TMP=$(mktemp -d)
SAN="-O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -std=c99 -D_GNU_SOURCE -w"
TYPED="davidson dav_ortho dav_project dav_restart dav_jdqmr dav_dynamic dav_refined front hostla"
for f in $ROOT/primme_b200/src/*.c; do gcc $SAN -c $f -o $TMP/$(basename $f .c).o & done; wait
for t in $TYPED; do gcc $SAN -DPB_COMPLEX -c $ROOT/primme_b200/src/$t.c -o $TMP/${t}_z.o & done; wait
gcc $SAN -shared -o $TMP/libprimme_hostcheck.so $TMP/*.o $ROOT/oracle/kernels_ref.c $ROOT/oracle/kernels_ref_z.c \
   -L$OB -l:libopenblasp-r0-59ffcd50.3.15.so -Wl,--disable-new-dtags,-rpath,$OB -lm
cp $ROOT/oracle/_build/libprimme_hostcheck.so $TMP/orig.so
trap "cp $TMP/orig.so $ROOT/oracle/_build/libprimme_hostcheck.so" EXIT
cp $TMP/libprimme_hostcheck.so $ROOT/oracle/_build/libprimme_hostcheck.so
TESTS=${@:-$ROOT/tests/test_jdqmr_cpu.py $ROOT/tests/test_zprimme_cpu.py $ROOT/tests/test_svds_cpu.py $ROOT/tests/test_host_logic.py $ROOT/tests/test_reference_golden_cpu.py $ROOT/tests/test_driver_cpu.py $ROOT/tests/test_refined_cpu.py $ROOT/tests/test_edge_cases_cpu.py $ROOT/tests/test_python_api.py}
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 \
   python -m pytest $TESTS -q -m "not gpu" -k "not crashes" 2>&1 | tee $TMP/out.txt | tail -3
echo "sanitizer reports: $(grep -c 'AddressSanitizer\|runtime error' $TMP/out.txt || true)"
grep -m5 'AddressSanitizer\|runtime error' $TMP/out.txt || true
