#include "../../ministark_b200/csrc/common.cuh"
#include "../../ministark_b200/csrc/ntt.cuh"
template __global__ void ms::k_ntt_fixed<ms::GL, 11, 2, 256, MS_NTT_MINB>(const ms::NttTile<ms::GL>);
