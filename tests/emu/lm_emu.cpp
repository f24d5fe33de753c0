// lm_emu.cpp -- pyro2_b200/csrc/lm.cu (kernels and host entry points, unchanged) compiled for the host through
// cuda_emu.h.  TEST INFRASTRUCTURE ONLY: built by tests/emu_util.py into tests/emu/liblm_emu.so (the p2b_lm_*
// C ABI over host memory).
#include <stdarg.h>

#include "cuda_emu_runtime.inc"

#include "../../pyro2_b200/csrc/lm.cu"

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

namespace {
struct RegisterThreaded {
    RegisterThreaded() { emu::threaded((const void*)pyro::lm_reduce_kernel); }
} register_threaded;
}  // namespace
