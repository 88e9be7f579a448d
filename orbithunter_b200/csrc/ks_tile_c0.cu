// tile path, symmetry class 0 (see ks_tile_host.cuh)
#define KS_TILE_CLS 0
#include "ks_tile_host.cuh"
