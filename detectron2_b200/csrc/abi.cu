// ABI bookkeeping entry points of libd2b200.so (include/d2b200.h).
#include "common.cuh"

D2B_API int d2b_abi_version(void) { return D2B_ABI_VERSION; }
D2B_API int d2b_cuda_version(void) { return CUDART_VERSION; }
D2B_API const char* d2b_arch(void) { return "sm_100a"; }
