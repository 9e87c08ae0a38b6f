// k_dtcwt_fwd.cu -- translation unit of dtcwt_fwd_stream.cuh (sm_100a)
#include "dtcwt_fwd_stream.cuh"
