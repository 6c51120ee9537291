// Interface between api.cu and final_fused.cu (the fused final-layer kernel's launcher).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "gemm_final.cuh"

namespace simstep {

struct FinalLaunch {
  CUtensorMap ax, ah, b;  // shared input rows, per-member activation rows, final-layer weights
  FinalArgs args;
};

// prec: SIMSTEP_PREC_*; args.groups in [1, 8].  Enqueues the launch on `st` (programmatic stream serialization).
cudaError_t launch_final_fused(int prec, const FinalLaunch& fl, int sm_count, int device, cudaStream_t st);

}  // namespace simstep
