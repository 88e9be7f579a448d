// 32x32 register-resident kernels, symmetry class 3 (see ks_w32_host.cuh)
#define KS_W32_CLS 3
#include "ks_w32_host.cuh"
