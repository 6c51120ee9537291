// Standalone bring-up test of simstep_debug_gemm (no torch): build with
//   nvcc -O2 -o tools/gemm_smoke tools/gemm_smoke.cu -Lamp_extensions_b200/csrc -lsimstep -Xlinker -rpath -Xlinker '$ORIGIN/../amp_extensions_b200/csrc'
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/simstep.h"

static float tf32_round(float x) {
  uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & ~0x1FFFu; memcpy(&x, &u, 4); return x;
}

int main(int argc, char** argv) {
  int prec = argc > 1 ? atoi(argv[1]) : 0;
  int groups = argc > 2 ? atoi(argv[2]) : 1;
  long long m = argc > 3 ? atoll(argv[3]) : 128;
  int n = argc > 4 ? atoi(argv[4]) : 256;
  int k = argc > 5 ? atoi(argv[5]) : 64;
  std::vector<float> a(size_t(groups) * m * k), b(size_t(groups) * n * k), d(size_t(groups) * m * n, NAN);
  srand(1);
  for (auto& v : a) v = tf32_round((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : b) v = tf32_round((rand() % 2001 - 1000) / 1000.f / sqrtf((float)k));
  if (prec != 0) { for (auto& v : a) v = roundf(v * 64) / 64; for (auto& v : b) v = roundf(v * 256) / 256; }
  float *ad, *bd, *dd;
  cudaMalloc(&ad, a.size() * 4); cudaMalloc(&bd, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
  cudaMemcpy(ad, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(bd, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dd, d.data(), d.size() * 4, cudaMemcpyHostToDevice);
  int rc = simstep_debug_gemm(prec, groups, m, n, k, ad, bd, nullptr, dd, nullptr);
  printf("rc=%d err=%s\n", rc, simstep_last_error(nullptr));
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  if (rc || e != cudaSuccess) return 1;
  cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; long long bad = 0;
  for (int g = 0; g < groups; ++g)
    for (long long i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int t = 0; t < k; ++t) s += double(a[(size_t(g) * m + i) * k + t]) * b[(size_t(g) * n + j) * k + t];
        double err = fabs(s - d[(size_t(g) * m + i) * n + j]);
        if (!(err <= 1e-4)) { if (bad < 10) printf("  bad g=%d i=%lld j=%d got=%g want=%g\n", g, i, j, d[(size_t(g) * m + i) * n + j], s); bad++; }
        if (err > maxerr) maxerr = err;
      }
  printf("max err %g, bad %lld of %zu\n", maxerr, bad, d.size());
  return bad ? 2 : 0;
}
