// tile path, symmetry class 3 (see ks_tile_host.cuh)
#define KS_TILE_CLS 3
#include "ks_tile_host.cuh"
