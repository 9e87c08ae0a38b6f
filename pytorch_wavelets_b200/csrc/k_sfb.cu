// k_sfb.cu -- translation unit of sfb_stream.cuh (sm_100a)
#include "sfb_stream.cuh"
