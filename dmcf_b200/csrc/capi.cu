// Library-wide pieces of the C ABI: version, thread-local error message, launch counter.
#include <cstdarg>

#include "common.cuh"

namespace dmcf {

std::atomic<int64_t> g_launches{0};

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace dmcf

extern "C" int dmcf_version(void) { return DMCF_B200_VERSION; }
extern "C" const char* dmcf_last_error(void) { return dmcf::error_buffer(); }
extern "C" int64_t dmcf_launch_count(void) { return dmcf::g_launches.load(std::memory_order_relaxed); }
