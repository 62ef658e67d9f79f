// R&D harness (not part of the product): loads the cubins written by gen.py and times them.
//   nvcc -O2 -std=c++17 tools/ptx_lab/run.cu -o tools/bin/ptx_lab_run -lcuda
//   tools/bin/ptx_lab_run tools/ptx_lab/out   (reads out/manifest.txt)
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#define CK(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s; cuGetErrorString(e, &s); printf("CUDA error %s at %s:%d\n", s, __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
  std::string dir = argc > 1 ? argv[1] : "tools/ptx_lab/out";
  CK(cuInit(0));
  CUdevice dev; CK(cuDeviceGet(&dev, 0));
  CUcontext ctx; CK(cuDevicePrimaryCtxRetain(&ctx, dev)); CK(cuCtxSetCurrent(ctx));
  int sms = 0, khz = 0;
  CK(cuDeviceGetAttribute(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
  CK(cuDeviceGetAttribute(&khz, CU_DEVICE_ATTRIBUTE_CLOCK_RATE, dev));
  printf("# %d SMs, max clock %d MHz; cycles are per SMSP per slice at max clock; 16 warps per SMSP\n", sms, khz / 1000);
  CUdeviceptr sink; CK(cuMemAlloc(&sink, 64));
  std::ifstream man(dir + "/manifest.txt");
  std::string line;
  CUevent e0, e1; CK(cuEventCreate(&e0, 0)); CK(cuEventCreate(&e1, 0));
  {  // warm the clocks: ~1.5 s of the first kernel
    std::ifstream m2(dir + "/manifest.txt"); std::string l2; std::getline(m2, l2);
    std::string first = l2.substr(0, l2.find('\t'));
    CUmodule mod; CK(cuModuleLoad(&mod, (dir + "/" + first + ".cubin").c_str()));
    CUfunction fn; CK(cuModuleGetFunction(&fn, mod, "k"));
    int iters = 16384; float seed = 1.0f; void* args[] = {&iters, &sink, &seed};
    for (int i = 0; i < 60; i++) CK(cuLaunchKernel(fn, sms * 8, 1, 1, 256, 1, 1, 0, 0, args, nullptr));
    CK(cuCtxSynchronize());
  }
  while (std::getline(man, line)) {
    std::stringstream ss(line);
    std::string name, ninstr, desc, hist;
    std::getline(ss, name, '\t'); std::getline(ss, ninstr, '\t'); std::getline(ss, desc, '\t'); std::getline(ss, hist, '\t');
    CUmodule mod; CK(cuModuleLoad(&mod, (dir + "/" + name + ".cubin").c_str()));
    CUfunction fn; CK(cuModuleGetFunction(&fn, mod, "k"));
    int iters = 16384; float seed = 1.0f;
    void* args[] = {&iters, &sink, &seed};
    const int blocks = sms * 8;  // 8 CTAs x 8 warps per SM = 16 warps per SMSP
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      CK(cuEventRecord(e0, 0));
      CK(cuLaunchKernel(fn, blocks, 1, 1, 256, 1, 1, 0, 0, args, nullptr));
      CK(cuEventRecord(e1, 0)); CK(cuEventSynchronize(e1));
      float ms; CK(cuEventElapsedTime(&ms, e0, e1));
      if (rep && ms < best) best = ms;
    }
    unsigned long long ck[3] = {0, 0, 0};
    CK(cuMemcpyDtoH(ck, sink, 24));
    const double mhz = ck[2] ? (double)ck[1] / (double)ck[2] * 1e3 : khz / 1e3;  // clock64 ticks per globaltimer ns
    const double slices_per_smsp = (double)blocks * 8 * iters * 4 / (sms * 4.0);
    const double cyc = best * 1e-3 * (mhz * 1e6) / slices_per_smsp;
    printf("%-18s %7.3f ms @%4.0f MHz %6.2f cyc/slice  (%5s instr/slice)  %s\n      SASS: %s\n", name.c_str(), best, mhz, cyc, ninstr.c_str(), desc.c_str(), hist.c_str());
    CK(cuModuleUnload(mod));
  }
  return 0;
}
