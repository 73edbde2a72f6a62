#include "abi_common.h"

namespace ara {

std::string& last_error_ref() {
    static thread_local std::string err;
    return err;
}

int set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return -1;
}

}  // namespace ara

extern "C" const char* ara_last_error(void) { return ara::last_error_ref().c_str(); }
