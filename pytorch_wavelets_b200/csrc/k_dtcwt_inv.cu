// k_dtcwt_inv.cu -- translation unit of dtcwt_inv_stream.cuh (sm_100a)
#include "dtcwt_inv_stream.cuh"
