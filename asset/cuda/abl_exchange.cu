// abl_exchange.cu — multi-GPU slab decomposition entry points (see abl_cuda.h).
// Placeholder: implemented after the single-GPU path is parity-green.
#include "abl_cuda.h"

extern "C" int abl_cuda_nccl_unique_id(void *) { return ABL_ERR_STATE; }
extern "C" int abl_cuda_comm_init_nccl(abl_runtime *, const void *, int, int) { return ABL_ERR_STATE; }
extern "C" int abl_cuda_set_slab(abl_runtime *, int, int) { return ABL_ERR_STATE; }
extern "C" int abl_cuda_slab_axis_layers(abl_runtime *, int *) { return ABL_ERR_STATE; }
extern "C" int abl_cuda_exchange(abl_runtime *, int) { return ABL_ERR_STATE; }
extern "C" int abl_cuda_owned_size(abl_runtime *, int, size_t *) { return ABL_ERR_STATE; }
