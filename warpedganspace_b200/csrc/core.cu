// Library-wide state: last-error string, launch counter, version / symbol sanity.
#include "common.cuh"
#include <string>

namespace wgs {

static thread_local std::string g_error;
unsigned long long g_launches = 0;

void set_error(const std::string& msg) { g_error = msg; }

int fail(const char* file, int line, const std::string& msg) {
    const char* base = file;
    for (const char* p = file; *p; ++p) if (*p == '/') base = p + 1;
    g_error = std::string(base) + ":" + std::to_string(line) + ": " + msg;
    return 1;
}

}  // namespace wgs

extern "C" const char* wgs_last_error() { return wgs::g_error.c_str(); }

extern "C" int wgs_version() { return 100; }   // 1.0.0

extern "C" unsigned long long wgs_launch_count() { return wgs::g_launches; }

extern "C" void wgs_reset_launch_count() { wgs::g_launches = 0; }

// Returns 0 when a usable sm_100 device is current, fills `sms` / `cc` when non-null.
extern "C" int wgs_device_info(int* sms, int* cc) {
    int dev = 0;
    WGS_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    WGS_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sms) *sms = prop.multiProcessorCount;
    if (cc) *cc = prop.major * 10 + prop.minor;
    WGS_REQUIRE(prop.major == 10, "libwgs_b200 is built for sm_100a only; found compute capability " +
                                      std::to_string(prop.major) + "." + std::to_string(prop.minor));
    return 0;
}
