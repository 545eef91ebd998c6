// Library info + error strings (host only).
#include "common.cuh"

#define SDB_STR2(x) #x
#define SDB_STR(x) SDB_STR2(x)

extern "C" int sdb_version(void) { return 100; }  // 0.1.0

extern "C" const char *sdb_build_info(void) {
    return "libsdb200 0.1.0 sm_100a nvcc " SDB_STR(__CUDACC_VER_MAJOR__) "." SDB_STR(__CUDACC_VER_MINOR__)
           " (tcgen05/TMEM fused render path, no CPU fallback)";
}

extern "C" const char *sdb_error_string(int code) {
    if (code == 0) return "success";
    if (code == SDB_EINVAL) return "sdb200: invalid argument";
    if (code == SDB_EUNSUPPORTED) return "sdb200: unsupported configuration";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "sdb200: unknown error";
}
