// bc_emu.cpp -- pyro2_b200/csrc/bc_user.cu compiled for the host through cuda_emu.h (TEST INFRASTRUCTURE
// ONLY; built by tests/emu_util.py into tests/emu/libbc_emu.so)
#include <stdarg.h>

#include "cuda_emu_runtime.inc"

#include "../../pyro2_b200/csrc/bc_user.cu"

namespace pyro {
char* last_error_buf()
{
    static thread_local char buf[512];
    return buf;
}
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
}
}  // namespace pyro

extern "C" const char* p2b_last_error(void) { return pyro::last_error_buf(); }
