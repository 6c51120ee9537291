// Launcher of the fused final-layer / env-step-tail kernel (gemm_final.cuh).  A translation unit of its own so that
// its (precision x member count) instantiations compile in parallel with api.cu.
#include <cuda.h>
#include <cuda_runtime.h>

#include "gemm_final.cuh"
#include "final_launch.h"

namespace simstep {

namespace {

template <typename E, int NM>
cudaError_t launch_t(const FinalLaunch& fl, int sm_count, int device, cudaStream_t st) {
  static size_t attr_smem[64] = {};  // function attributes are per device
  auto kern = final_fused_kernel<E, NM>;
  const size_t smem = final_smem_bytes(fl.args.S, fl.args.groups);
  size_t& have = attr_smem[device % 64];
  if (have < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    have = smem;
  }
  const int total = fl.args.m_tiles * fl.args.groups;
  if (total <= 0) return cudaSuccess;
  const int slots = total < sm_count / 2 ? total : sm_count / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(slots * 2));
  cfg.blockDim = dim3(kFinalThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kern, fl.ax, fl.ah, fl.b, fl.args);
}

template <typename E>
cudaError_t launch_n(const FinalLaunch& fl, int sm_count, int device, cudaStream_t st) {
  switch (fl.args.groups) {
    case 1: return launch_t<E, 1>(fl, sm_count, device, st);
    case 2: return launch_t<E, 2>(fl, sm_count, device, st);
    case 3: return launch_t<E, 3>(fl, sm_count, device, st);
    case 4: return launch_t<E, 4>(fl, sm_count, device, st);
    case 5: return launch_t<E, 5>(fl, sm_count, device, st);
    case 6: return launch_t<E, 6>(fl, sm_count, device, st);
    case 7: return launch_t<E, 7>(fl, sm_count, device, st);
    case 8: return launch_t<E, 8>(fl, sm_count, device, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_final_fused(int prec, const FinalLaunch& fl, int sm_count, int device, cudaStream_t st) {
  switch (prec) {
    case SIMSTEP_PREC_TF32: return launch_n<ElemTF32>(fl, sm_count, device, st);
    case SIMSTEP_PREC_FP16: return launch_n<ElemF16>(fl, sm_count, device, st);
    case SIMSTEP_PREC_BF16: return launch_n<ElemBF16>(fl, sm_count, device, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace simstep
