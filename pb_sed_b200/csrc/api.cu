// api.cu -- library identification and launch accounting
#include "common.cuh"

long long g_pbsed_launches = 0;
const char* g_pbsed_last_kernel = "";

extern "C" int pbsed_abi_version(void) { return 6; }
extern "C" long long pbsed_launch_count(void) { return g_pbsed_launches; }
extern "C" const char* pbsed_last_kernel(void) { return g_pbsed_last_kernel; }
