// Specialisation of the chain-batched kernels for the MLP 2321 architecture (fp32 + fp64).
#include "inst_common.cuh"
EB_INSTANTIATE_NET(2321, LOSS_BINARY, 2, 3, 2, 1)
