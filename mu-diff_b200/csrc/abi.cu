// ABI bookkeeping for libmudiff_b200.so
#include "common.cuh"

int64_t g_mudiff_launches = 0;

extern "C" int mudiff_abi_version(void) { return 1; }

extern "C" const char* mudiff_build_info(void) {
  return "libmudiff_b200 abi=1 arch=sm_100a cuda=" 
#define STR2(x) #x
#define STR(x) STR2(x)
      STR(__CUDACC_VER_MAJOR__) "." STR(__CUDACC_VER_MINOR__) " built " __DATE__;
}

extern "C" int64_t mudiff_launch_count(void) { return g_mudiff_launches; }

extern "C" int mudiff_conv_desc_size(void) { return (int)sizeof(mudiff_conv_desc); }
