// Library identification + device queries shared by all kernels.
#include <cstdio>
#include <cstring>

#include "common.cuh"

int bts_num_sms() {
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        n = 148;   // B200
    cached = n;
    return n;
}

extern "C" int bts_version(char *buf, int buflen) {
    static const char v[] = "bts_b200 0.1 (sm_100a; lpg, plane_head, silog, conv_tc)";
    if (buf && buflen > 0) {
        std::strncpy(buf, v, (size_t)buflen - 1);
        buf[buflen - 1] = 0;
    }
    return 100;
}
