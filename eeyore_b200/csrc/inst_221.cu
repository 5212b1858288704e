// Specialisation of the chain-batched kernels for the MLP 221 architecture (fp32 + fp64).
#include "inst_common.cuh"
EB_INSTANTIATE_NET(221, LOSS_BINARY, 2, 2, 1)
