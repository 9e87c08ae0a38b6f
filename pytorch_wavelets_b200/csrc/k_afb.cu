// k_afb.cu -- translation unit of afb_stream.cuh (sm_100a)
#include "afb_stream.cuh"
