// k = 5..8 instantiations (KD = 8) of the warp-per-candidate likelihood kernel: see krig_nll_small.cu
#define SMB_SMALL_K8_TU
#include "krig_nll_small.cu"
