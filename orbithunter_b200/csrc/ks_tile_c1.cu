// tile path, symmetry class 1 (see ks_tile_host.cuh)
#define KS_TILE_CLS 1
#include "ks_tile_host.cuh"
