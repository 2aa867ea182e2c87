#!/bin/bash
# Host-side memory / UB check of the engine sources: builds the host-compiled kernel build (tests/emu) with
# AddressSanitizer + UBSan and runs the CPU parity, error-path, checkpoint and multi-rank (gloo) tests under it.
# The CUDA library shares hk_engine.cu (all host plumbing) and the kernel bodies with this build.
#   bash scripts/asan_check.sh        -> prints the number of sanitizer reports (0 expected) and the pytest tail
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
cd "$ROOT/hakai_fem_b200/csrc" || exit 1
ASAN_LIB="$(/usr/bin/g++ -print-file-name=libasan.so)"
/usr/bin/g++ -x c++ -DHK_EMU -O1 -g -std=c++17 -fPIC -ffp-contract=off -fsanitize=address,undefined \
    -fno-omit-frame-pointer -Wno-unknown-pragmas -shared -o ../../tests/emu/libhakai_emu.so \
    hk_exact.cu hk_element.cu hk_engine.cu hk_setup.cu || exit 1
cd "$ROOT" || exit 1
ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 LD_PRELOAD="$ASAN_LIB" \
    python -m pytest tests/test_emu_parity.py tests/test_abi_errors.py tests/test_checkpoint.py tests/test_multi_gloo.py \
    -q -s -p no:cacheprovider > /tmp/hk_asan.log 2>&1
echo "sanitizer reports: $(grep -c 'AddressSanitizer\|runtime error' /tmp/hk_asan.log)"
tail -n 2 /tmp/hk_asan.log
rm -f tests/emu/libhakai_emu.so && make -s -C hakai_fem_b200/csrc emu     # back to the optimised build
