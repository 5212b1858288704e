// Specialisation of the chain-batched kernels for the MLP 433 architecture (fp32 + fp64).
#include "inst_common.cuh"
EB_INSTANTIATE_NET(433, LOSS_MULTICLASS, 4, 3, 3)
