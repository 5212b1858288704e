// Specialisation of the chain-batched kernels for the MLP 4323 architecture (fp32 + fp64).
#include "inst_common.cuh"
EB_INSTANTIATE_NET(4323, LOSS_MULTICLASS, 4, 3, 2, 3)
