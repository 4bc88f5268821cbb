#!/bin/bash
# The CPU test suite against ASan/UBSan builds of the oracle and of the library's host code (the device code is compiled
# as usual; without a GPU only host paths run: relation validation, skip lists, writers, plan and source generation,
# numeric formatting, the exchange plans).  Restores the normal builds afterwards.
#   bash tools/sanitize_host.sh
# Round 2: 51 tests, no sanitizer report.
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
ASAN=$(gcc -print-file-name=libasan.so)
TMP=$(mktemp -d)
python -c "from citus_b200 import build; build.build()"; make -s -C oracle liboracle.so
cp citus_b200/lib/libcitus_gpu.so "$TMP/libcitus_gpu.so.bak"; cp oracle/liboracle.so "$TMP/liboracle.so.bak"
restore() { cp "$TMP/libcitus_gpu.so.bak" citus_b200/lib/libcitus_gpu.so; cp "$TMP/liboracle.so.bak" oracle/liboracle.so; touch citus_b200/lib/libcitus_gpu.so oracle/liboracle.so; }
trap restore EXIT
gcc -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -shared -fPIC -o oracle/liboracle.so oracle/oracle.c -ldl -lpthread -lm
SRCS=$(python -c "from citus_b200 import build; print(' '.join(build.SOURCES))")
for s in $SRCS; do
	/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 \
		-Xcompiler -fPIC,-fopenmp,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer \
		-I include -I citus_b200/csrc -x cu -c citus_b200/csrc/$s -o "$TMP/$s.o" 2> /dev/null &
done
wait
/usr/local/cuda/bin/nvcc -shared -o citus_b200/lib/libcitus_gpu.so "$TMP"/*.o -Xcompiler -fopenmp,-fsanitize=address,-fsanitize=undefined -lgomp -ldl 2> /dev/null
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 python -m pytest tests/test_host_cabi.py tests/test_distributed_gloo.py tests/test_oracle_golden.py -x -q -m "not gpu"
