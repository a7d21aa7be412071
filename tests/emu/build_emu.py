"""TEST INFRASTRUCTURE: builds tests/emu/libps21_emu.so = the general path of 21cmmc_b200/csrc
compiled by g++ against the host CUDA emulation (cuda_emu.h).  Used only by `-m "not gpu"` tests."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "21cmmc_b200", "csrc")
LIB = os.path.join(HERE, "libps21_emu.so")


def build(force=False):
    if os.environ.get("PS21_EMU_LIBRARY"):   # e.g. an AddressSanitizer build of the same sources
        return os.environ["PS21_EMU_LIBRARY"]
    srcs = [os.path.join(CSRC, f) for f in ("ps21_plan.cu", "ps21_general.cu", "ps21_fused.cu", "ps21_g2.cu")]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "cuda_emu.*")) + [
        os.path.join(ROOT, "include", "ps21.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-DPS21_EMU", "-ffp-contract=off", "-fPIC", "-shared", "-w",
           "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-o", LIB,
           os.path.join(HERE, "cuda_emu.cpp")]
    for s in srcs:
        cmd += ["-x", "c++", s]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
