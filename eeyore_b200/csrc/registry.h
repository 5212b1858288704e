// Type-erased table of the compiled network specialisations (one entry per architecture x dtype).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/eeyore_b200.h"

namespace eb {

enum { KIND_MH = 0, KIND_MALA = 1, KIND_HMC = 2, KIND_HMC_TUNED = 3 };  // TUNED: HMC + per-chain dual averaging

struct EvalCall {
  int64_t n_chains;
  const void *theta, *x, *y;
  int64_t n_rows;
  const void *ploc, *pscale;
  int has_temperature;
  double temperature;
  void *out_target, *out_grad, *out_ll, *out_lp;
  int lanes;
  int use_bulk;
  cudaStream_t stream;
};

struct NetEntry {
  int n_layers;
  int dims[4];
  int loss;
  int dtype;
  int n_params;
  cudaError_t (*eval)(const EvalCall&);
  cudaError_t (*sampler)(int kind, const eeyore_b200_run_params&, int lanes, int use_bulk);
  cudaError_t (*forward)(int64_t n_chains, const void* theta, const void* x, int64_t n_rows, void* out, cudaStream_t);
  cudaError_t (*smmala)(const eeyore_b200_run_params&, int use_bulk);
  cudaError_t (*adaptive)(int kind, const eeyore_b200_run_params&, int use_bulk);   // AM / RAM (P <= 32), else nullptr
};

}  // namespace eb
