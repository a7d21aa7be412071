# Bounds check of the kernels on the CPU: the emulated build (tests/emu) compiled with AddressSanitizer, then the
# emulated parity tests run against it.  (compute-sanitizer is not available on the GPU pool.)
set -e
mkdir -p build/asan
g++ -O1 -g -std=c++17 -DPS21_EMU -ffp-contract=off -fPIC -shared -w -fsanitize=address -fno-omit-frame-pointer \
    -I tests/emu -I 21cmmc_b200/csrc -I include -o build/asan/libps21_emu_asan.so tests/emu/cuda_emu.cpp \
    -x c++ 21cmmc_b200/csrc/ps21_plan.cu -x c++ 21cmmc_b200/csrc/ps21_general.cu -x c++ 21cmmc_b200/csrc/ps21_fused.cu
PS21_EMU_LIBRARY=$PWD/build/asan/libps21_emu_asan.so LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
python -m pytest tests/test_emu_parity.py tests/test_property_sweep.py tests/test_plugin_api.py -x -q
