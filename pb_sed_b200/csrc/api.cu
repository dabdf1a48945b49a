// api.cu -- library identification and launch accounting
#include "common.cuh"

long long g_pbsed_launches = 0;

extern "C" int pbsed_abi_version(void) { return 5; }
extern "C" long long pbsed_launch_count(void) { return g_pbsed_launches; }
