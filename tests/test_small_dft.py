"""The register butterflies of 21cmmc_b200/csrc/fft_smem.cuh (scalar forms) against a direct DFT, on the host."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#define PS21_EMU
#include "fft_smem.cuh"
#include <complex>
#include <cstdio>
namespace ps21emu { uint3 threadIdx, blockIdx; dim3 blockDim, gridDim; char* g_dyn_smem; void syncthreads(){} void syncwarp(){}
  void spin_yield(){} unsigned long long* shfl_slot(int){return nullptr;} int linear_tid(){return 0;}
  void launch(dim3,dim3,size_t,const std::function<void()>&){} }
template<int N> double check() {
  float2 v[N]; std::complex<double> x[N];
  for (int i=0;i<N;++i){ v[i]=make_float2((float)std::sin(1.3*i+0.2),(float)std::cos(0.7*i*i)); x[i]={v[i].x,v[i].y}; }
  ps21::small_dft<N>(v);
  double err=0;
  for (int k=0;k<N;++k){ std::complex<double> s=0; for(int i=0;i<N;++i) s+=x[i]*std::polar(1.0,-2*M_PI*i*k/N);
    err=std::max(err,std::abs(s-std::complex<double>(v[k].x,v[k].y))); }
  return err;
}
int main(){ printf("%g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g\n",check<2>(),check<3>(),check<4>(),check<5>(),check<8>(),check<16>(),check<32>(),check<25>(),check<7>(),check<11>(),check<13>(),check<6>(),check<9>(),check<10>(),check<12>(),check<15>()); }
'''


def test_small_dfts_match_direct_dft(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "tests", "emu"), "-I",
                    os.path.join(ROOT, "21cmmc_b200", "csrc"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    errs = [float(x) for x in out]
    assert len(errs) == 16
    # inputs are O(1); a 32-point fp32 transform accumulates ~1e-6 absolute error at most
    assert max(errs) < 3e-6, errs
