# AddressSanitizer build of the host emulation (tests/emul) for tools/emul_asan_check.py and tools/emul_asan_sweep.py:
#   bash tools/build_emul_asan.sh && LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/emul_asan_sweep.py 150 3
set -e
cd "$(dirname "$0")/.."
mkdir -p tests/emul/_build/asan
FL="-x c++ -std=c++17 -O1 -g -fPIC -pthread -fsanitize=address -fno-omit-frame-pointer -DSFC_HOST_EMUL -Itests/emul -Iscirs_b200/csrc -Iinclude"
for f in plan aux_kernels kernels_f64_small kernels_f64_mid kernels_f64_big kernels_f64_real kernels_f64_dbl_a kernels_f64_dbl_b kernels_dct; do
  g++ $FL -c scirs_b200/csrc/$f.cu -o tests/emul/_build/asan/$f.o &
done
g++ $FL -c tests/emul/plan_emul.cpp -o tests/emul/_build/asan/plan_emul.o &
wait
g++ -shared -pthread -fsanitize=address -o tests/emul/_build/asan/libplan_emul.so tests/emul/_build/asan/*.o
echo built tests/emul/_build/asan/libplan_emul.so
