// TEST INFRASTRUCTURE: the reference's WHOLE program (main.cpp + APD.cpp + APD.cu, all unmodified and included from
// /root/reference where they lie) as one executable, oracle/_ref/apd_main_ref, against oracle/shim_host. One seam:
// clock64() -> a constant, so that the curand seed (APD.cu:803) is reproducible and equals the APD_SEED the facade build
// (oracle/_ref/apd_main_b200) is run with. Build: oracle/Makefile target `main`.
#include <opencv2/opencv.hpp>
#include <boost/filesystem.hpp>
#include <cuda_runtime.h>
#include <cuda.h>
#include <curand_kernel.h>
#define clock64() (1234567ULL)
#include APD_REF_CU
#include APD_REF_CPP
#include APD_REF_MAIN
