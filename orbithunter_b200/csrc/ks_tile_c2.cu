// tile path, symmetry class 2 (see ks_tile_host.cuh)
#define KS_TILE_CLS 2
#include "ks_tile_host.cuh"
